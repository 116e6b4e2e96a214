#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for s in 6 8 12; do
  timeout 300 python bench.py --no-cpu-baseline --steps 400 --in-flight $s 2>/dev/null | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('b1 in_flight $s value %.0f e2e %.0f frac %.3f'%(r['value'], r['e2e']['value'], r['roofline']['frac']))"
done
timeout 300 python bench.py --no-cpu-baseline --steps 80 --batch 8 --in-flight 4 2>/dev/null | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('b8 in_flight 4 value %.0f e2e %.0f frac %.3f'%(r['value'], r['e2e']['value'], r['roofline']['frac']))"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
