#!/bin/bash
# quick iteration: GPU tests (optionally a -k filter in $1), SM-time table, default detect throughput
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -x ${1:+-k "$1"} > gpurun_out/pytest_iter.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_iter.log
bash tools/gpu_smtime.sh b1 | tail -28
python bench.py --workload detect --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('detect us_per_frame %.1f  img/s %.0f  e2e %.0f  sync_us %.1f frac %.3f' % (1e3*d['ms_per_step'], d['value'], d['e2e']['value'], 1e3*d['config']['sync']['ms_per_step'], d['roofline']['frac']))"
