#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_nms.py tests/test_golden_next_rows.py -x -q -m gpu 2>&1 | tail -2
for n in 4000 1000000; do timeout 200 python bench.py --workload nms --nms-n $n --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('nms $n value %.3g e2e %.3g d2h %d'%(r['value'], r['e2e']['value'], r['e2e']['d2h_bytes_per_step']))"; done
