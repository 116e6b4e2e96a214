#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_labelling.py tests/test_optim.py -x -q -m gpu > gpurun_out/pytest_new.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_new.log | cut -c1-220
