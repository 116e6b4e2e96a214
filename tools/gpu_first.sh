#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -k "first_layer or pnet or vgg_large" > gpurun_out/pytest_first.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_first.log
FRCNN_FIRST_TMA=0 timeout 200 python tools/bench_conv_first.py
timeout 200 python tools/bench_conv_first.py
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
timeout 300 python bench.py --no-cpu-baseline --steps 300 > gpurun_out/bench_b1.log 2>&1; show gpurun_out/bench_b1.log "b1 if6"
timeout 300 python bench.py --no-cpu-baseline --steps 60 --batch 8 --in-flight 3 > gpurun_out/bench_b8.log 2>&1; show gpurun_out/bench_b8.log "b8 if3"
