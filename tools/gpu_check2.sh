#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_detect.py tests/test_gpu_conv_backward.py -x -q > gpurun_out/pytest_t2.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_t2.log | cut -c1-200
timeout 300 python bench.py --workload train --batch 8 --steps 5 --warmup 3 > gpurun_out/bench_train.log 2>&1; tail -1 gpurun_out/bench_train.log | cut -c1-220
timeout 300 python bench.py --no-cpu-baseline --steps 300 > gpurun_out/bench_b1.log 2>&1; tail -1 gpurun_out/bench_b1.log | cut -c1-160
bash tools/gpu_trainll.sh 2>&1 | tail -28
