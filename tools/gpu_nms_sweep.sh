#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/r1b_nms_sweep.jsonl
for n in 1000 4000 16000 64000 256000 1000000; do
  timeout 300 python bench.py --workload nms --nms-n $n --steps 10 --warmup 3 2>/dev/null | tail -1 >> gpurun_out/r1b_nms_sweep.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r1b_nms_sweep.jsonl'):
    r=json.loads(l)
    print(r['config']['workload'][:40], 'value %.3g e2e %.3g cpu %.3g'%(r['value'], r['e2e']['value'], r.get('cpu_baseline',{}).get('value',0)))
PY
