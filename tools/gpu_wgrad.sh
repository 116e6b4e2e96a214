#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_conv_backward.py -x -q > gpurun_out/pytest_wg.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_wg.log | cut -c1-220
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_nms.py -x -q > gpurun_out/pytest_wg2.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_wg2.log | cut -c1-220
FRCNN_WGRAD_HALO=0 timeout 300 python bench.py --workload train --batch 8 --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-140
timeout 300 python bench.py --workload train --batch 8 --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-140
timeout 120 python bench.py --workload nms --nms-n 4000 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('nms 4000 value %.3g e2e %.3g'%(r['value'], r['e2e']['value']))"
