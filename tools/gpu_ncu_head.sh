#!/bin/bash
# ncu --set full of the fused anchor-network kernel (one launch of a synchronous batch-1 step)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
K=${1:-conv_head_kernel}
FRCNN_NO_GRAPH=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$K" -s 3 -c 1 -o /tmp/r2_head -f python bench.py --workload detect --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 > gpurun_out/ncu_head.log 2>&1
ncu -i /tmp/r2_head.ncu-rep --page raw --csv > gpurun_out/r2_head_raw.csv 2>/dev/null
ncu -i /tmp/r2_head.ncu-rep --page details > gpurun_out/r2_head_details.txt 2>/dev/null
ncu -i /tmp/r2_head.ncu-rep --page source --csv > gpurun_out/r2_head_source.csv 2>/dev/null
ls -la /tmp/r2_head.ncu-rep gpurun_out/r2_head_*; tail -3 gpurun_out/ncu_head.log
