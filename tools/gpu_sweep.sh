#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
FRCNN_BENCH_CFG="256,12;256,11;192,12;192,11;128,12;128,11;256,1;192,1;128,2;128,1" timeout 600 python tools/bench_conv_layers.py 1 8 > gpurun_out/sweep.log 2>&1
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/sweep.log') if l.startswith('{')]
from collections import defaultdict
d=defaultdict(list)
for r in rows: d[(r['batch'],r['layer'])].append((r['us'],tuple(r['cfg'])))
for k,v in d.items(): print(k, sorted(v))
PY
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_b1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b1.log 2>&1
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_b8.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --batch 8 > gpurun_out/ncu_b8.log 2>&1
python tools/parse_launches.py gpurun_out/launches_b1.csv 2>&1 | tail -40
python tools/parse_launches.py gpurun_out/launches_b8.csv 2>&1 | tail -40
