#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log
timeout 300 python bench.py --workload nms > gpurun_out/bench_nms.log 2>&1; tail -1 gpurun_out/bench_nms.log | cut -c1-400
timeout 300 python bench.py --workload train --batch 8 --steps 5 --warmup 3 > gpurun_out/bench_train.log 2>&1; tail -1 gpurun_out/bench_train.log | cut -c1-400
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-400
