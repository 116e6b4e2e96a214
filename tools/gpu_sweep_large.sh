#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
FRCNN_BENCH_MODEL=large FRCNN_BENCH_CFG="64,12;64,11;64,2;64,1;128,12;128,2;256,12;256,11;256,1" timeout 600 python tools/bench_conv_layers.py 1 > gpurun_out/sweep_large.log 2>&1
python - <<'PY'
import json
from collections import defaultdict
d=defaultdict(list)
for l in open('gpurun_out/sweep_large.log'):
    if l.startswith('{'):
        r=json.loads(l); d[(r['batch'],r['layer'])].append((r['us'],tuple(r['cfg']),r['tflops']))
for k,v in d.items(): print(k, sorted(v))
PY
tail -3 gpurun_out/sweep_large.log | cut -c1-200
