#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "swapped" 2>&1 | tail -6
for l in conv2_1 "conv2_2+pool" conv4_1 "conv4_2+pool"; do
  FRCNN_BENCH_LAYER=$l FRCNN_BENCH_CFG="0,61;128,31;192,12" python tools/bench_conv_layers.py 1 8 2>&1 | tail -6
done
for cfg in "conv2_1 0,61" "conv2_2+pool 0,61" "conv4_2+pool 0,61"; do
  set -- $cfg
  echo "== $1 cfg $2"
  FRCNN_CONV_TRACE=gpurun_out/conv_trace.bin FRCNN_BENCH_LAYER=$1 FRCNN_BENCH_CFG="$2" python tools/bench_conv_layers.py 1 2>&1 | tail -1
  python tools/conv_trace.py gpurun_out/conv_trace.bin 2>/dev/null | head -7
done
