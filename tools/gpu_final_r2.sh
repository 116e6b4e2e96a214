#!/bin/bash
# final evidence of the round on one GPU: whole GPU suite, smoke, the default bench line (all BASELINE configurations),
# the training line alone and its launch list
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 300 gpurun_out/r2_bench_default.json
timeout 300 python bench.py --workload train --steps 20 --warmup 5 > gpurun_out/r2_bench_train_1gpu.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print('detect', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac %.3f' % d['roofline']['frac'], d['clocks'])
for k,v in d.get('also',{}).items(): print(k, round(v['value']), v['unit'], 'e2e', round(v['e2e']['value']), 'frac', v.get('roofline',{}).get('frac'))
v=json.loads(open('gpurun_out/r2_bench_train_1gpu.json').read().strip().splitlines()[-1])
print('train alone', round(v['value']), 'e2e', round(v['e2e']['value']), 'ms/step %.3f' % v['ms_per_step'], 'frac %.3f' % v['roofline']['frac'], 'cpu', v.get('cpu_baseline',{}).get('value'))
PY
bash tools/gpu_trainll.sh > gpurun_out/r2_launches_train_b8.txt 2>&1; head -40 gpurun_out/r2_launches_train_b8.txt
