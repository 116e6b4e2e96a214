#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -k "first_layer" 2>&1 | tail -2
timeout 200 python tools/bench_conv_first.py 2>&1 | tail -2
for b in 1 8; do
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 260 --csv --log-file gpurun_out/launches_thr_b$b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 --batch $b > gpurun_out/ncu_thr_b$b.log 2>&1
python tools/parse_launches.py gpurun_out/launches_thr_b$b.csv 2>&1 | tail -30
done
