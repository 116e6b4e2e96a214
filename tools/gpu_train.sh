#!/bin/bash
# training path check: parity tests of the backward primitives / objective, then the 8-frame training bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_conv_backward.py tests/test_optim.py -x -q -m gpu > gpurun_out/pytest_train.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_train.log | cut -c1-220
timeout 300 python bench.py --workload train --batch 8 --steps 5 --warmup 3 > gpurun_out/bench_train.log 2>&1; tail -1 gpurun_out/bench_train.log | cut -c1-140
