#!/bin/bash
# SM-time table of one synchronous batch-1 detect step: ncu launch list with the average SM-active cycles of every launch
# (= mean CTA residency for the one-CTA-per-SM kernels: what a frame costs when other frames fill the idle SMs)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg --clock-control none -c 300 --csv --log-file gpurun_out/smtime_${1:-b1}.csv python bench.py --workload detect --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 ${@:2} > /dev/null 2>&1
python tools/smtime.py gpurun_out/smtime_${1:-b1}.csv
