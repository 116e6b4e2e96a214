#!/bin/bash
# GPU session 1: settle the halo-kernel descriptor mode, then parity + per-layer timings + bench.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="tests/test_gpu_conv.py::test_conv_exact_integers tests/test_gpu_conv.py::test_conv_halo_kernel tests/test_gpu_conv.py::test_conv_halo_equals_tap_kernel"
FRCNN_HALO_DESC=0 timeout 600 python -m pytest $T -x -q > gpurun_out/halo_desc0.log 2>&1
rc0=$?
echo "desc0 rc=$rc0"; tail -5 gpurun_out/halo_desc0.log
if [ $rc0 -ne 0 ]; then
  FRCNN_HALO_DESC=1 timeout 600 python -m pytest $T -x -q > gpurun_out/halo_desc1.log 2>&1
  rc1=$?
  echo "desc1 rc=$rc1"; tail -5 gpurun_out/halo_desc1.log
  if [ $rc1 -eq 0 ]; then export FRCNN_HALO_DESC=1; else export FRCNN_CONV_HALO=0; fi
fi
echo "mode: HALO_DESC=${FRCNN_HALO_DESC:-0} CONV_HALO=${FRCNN_CONV_HALO:-1}" | tee gpurun_out/halo_mode.txt
FRCNN_CONV_HALO=0 timeout 300 python tools/bench_conv_layers.py > gpurun_out/layers_tap.log 2>&1; tail -1 gpurun_out/layers_tap.log
timeout 300 python tools/bench_conv_layers.py > gpurun_out/layers_halo.log 2>&1; tail -1 gpurun_out/layers_halo.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_b1.log 2>&1; tail -1 gpurun_out/bench_b1.log
FRCNN_CONV_HALO=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_b1_tap.log 2>&1; tail -1 gpurun_out/bench_b1_tap.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --batch 8 > gpurun_out/bench_b8.log 2>&1; tail -1 gpurun_out/bench_b8.log | cut -c1-300
