#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/bench_conv_first.py 1 > gpurun_out/san_first.log 2>&1
grep -v "^$" gpurun_out/san_first.log | head -60
