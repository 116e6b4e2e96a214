#!/bin/bash
# launch list of one vgg_large 1000x600 detect step (BASELINE configs[3], per GPU)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file gpurun_out/launches_r1b_large_b1.csv python bench.py --model vgg_large --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 > /dev/null 2>&1
python tools/parse_launches.py gpurun_out/launches_r1b_large_b1.csv 2>&1 | tail -36
