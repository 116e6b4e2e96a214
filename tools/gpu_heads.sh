#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_detect.py -x -q > gpurun_out/pytest_part.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_part.log
show() { python - "$1" "$2" <<'PY'
import json, sys
f, tag = sys.argv[1], sys.argv[2]
try:
    r=json.loads(open(f).read().strip().splitlines()[-1])
    print(tag, 'value %.0f e2e %.0f sync %.0f frac %.3f sync_frac %.3f'%(r['value'], r['e2e']['value'], r['config']['sync']['images_per_sec'], r['roofline']['frac'], r['roofline']['sync_pass']['frac']))
except Exception as e:
    print(tag, 'fail', e); print(open(f).read()[-2000:])
PY
}
for sch in latency throughput; do
  timeout 300 python bench.py --no-cpu-baseline --steps 300 --in-flight 6 --schedule $sch > gpurun_out/bench_b1_$sch.log 2>&1; show gpurun_out/bench_b1_$sch.log "b1 if6 $sch"
  timeout 300 python bench.py --no-cpu-baseline --steps 60 --batch 8 --in-flight 3 --schedule $sch > gpurun_out/bench_b8_$sch.log 2>&1; show gpurun_out/bench_b8_$sch.log "b8 if3 $sch"
done
timeout 300 python bench.py --no-cpu-baseline --steps 300 --in-flight 8 > gpurun_out/bench_b1_if8.log 2>&1; show gpurun_out/bench_b1_if8.log "b1 if8 throughput"
