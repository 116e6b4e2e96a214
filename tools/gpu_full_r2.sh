#!/bin/bash
# everything once: the whole GPU suite, smoke, the train / detect bench lines
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac %.3f' % d['roofline']['frac'])"
python bench.py --workload detect --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('detect us_per_frame %.1f  img/s %.0f  e2e %.0f  sync_us %.1f frac %.3f' % (1e3*d['ms_per_step'], d['value'], d['e2e']['value'], 1e3*d['config']['sync']['ms_per_step'], d['roofline']['frac']))"
python bench.py --workload detect --model vgg_large --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('vgg_large img/s %.0f  e2e %.0f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']))"
bash tools/gpu_smtime.sh b1 | tail -27
