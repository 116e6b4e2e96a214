#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_detect.py tests/test_gpu_conv.py -x -q > gpurun_out/pytest_part.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_part.log
for s in 1 2 3 4; do
  timeout 300 python bench.py --no-cpu-baseline --steps 200 --in-flight $s > gpurun_out/bench_b1_if$s.log 2>&1; python - <<PY
import json
try:
    r=json.loads(open('gpurun_out/bench_b1_if$s.log').read().strip().splitlines()[-1])
    print('b1 in_flight', $s, 'value %.0f e2e %.0f sync %.0f frac %.3f'%(r['value'], r['e2e']['value'], r['config']['sync']['images_per_sec'], r['roofline']['frac']))
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_b1_if$s.log').read()[-1500:])
PY
done
for s in 1 2; do
  timeout 300 python bench.py --no-cpu-baseline --steps 50 --batch 8 --in-flight $s > gpurun_out/bench_b8_if$s.log 2>&1; python - <<PY
import json
try:
    r=json.loads(open('gpurun_out/bench_b8_if$s.log').read().strip().splitlines()[-1])
    print('b8 in_flight', $s, 'value %.0f e2e %.0f sync %.0f frac %.3f'%(r['value'], r['e2e']['value'], r['config']['sync']['images_per_sec'], r['roofline']['frac']))
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_b8_if$s.log').read()[-1500:])
PY
done
