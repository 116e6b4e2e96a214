#!/bin/bash
mkdir -p gpurun_out
FRCNN_NO_GRAPH=1 FRCNN_HEAD_TRACE=gpurun_out/head_trace.bin timeout 300 python bench.py --workload detect --steps 3 --warmup 1 --no-cpu-baseline --in-flight 1 > /dev/null 2>gpurun_out/head_trace.err
python tools/head_trace.py gpurun_out/head_trace.bin 14
