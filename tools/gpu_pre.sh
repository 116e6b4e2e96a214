#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_preprocess.py -x -q -m gpu > gpurun_out/pytest_pre.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_pre.log | cut -c1-220
timeout 300 python tools/bench_next_rows.py > gpurun_out/r1b_next_rows.jsonl 2>&1; cut -c1-330 gpurun_out/r1b_next_rows.jsonl
