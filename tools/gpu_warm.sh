#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/launches_b1_warm.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 > gpurun_out/ncu_b1w.log 2>&1
python tools/parse_launches.py gpurun_out/launches_b1_warm.csv 2>&1 | tail -30
for s in 6 8; do
  timeout 300 python bench.py --no-cpu-baseline --steps 300 --in-flight $s > gpurun_out/bench_b1_if$s.log 2>&1; python - <<PY
import json
try:
    r=json.loads(open('gpurun_out/bench_b1_if$s.log').read().strip().splitlines()[-1])
    print('b1 in_flight', $s, 'value %.0f e2e %.0f sync %.0f frac %.3f'%(r['value'], r['e2e']['value'], r['config']['sync']['images_per_sec'], r['roofline']['frac']))
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_b1_if$s.log').read()[-1500:])
PY
done
