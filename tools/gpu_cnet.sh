#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for ctas in 148 74 37 16 8; do
  FRCNN_CNET_CTAS=$ctas timeout 300 python bench.py --no-cpu-baseline --steps 300 --in-flight 6 > gpurun_out/bench_cnet$ctas.log 2>&1; python - <<PY
import json
try:
    r=json.loads(open('gpurun_out/bench_cnet$ctas.log').read().strip().splitlines()[-1])
    print('cnet ctas', $ctas, 'value %.0f e2e %.0f sync %.0f'%(r['value'], r['e2e']['value'], r['config']['sync']['images_per_sec']), r['roofline']['stage_ms_per_step'])
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_cnet$ctas.log').read()[-1500:])
PY
done
FRCNN_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_first -s 3 -c 1 -o gpurun_out/conv_first_b1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --in-flight 1 > gpurun_out/ncu_first.log 2>&1
FRCNN_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_igemm -s 8 -c 2 -o gpurun_out/conv_tap_b1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --in-flight 1 > gpurun_out/ncu_tap.log 2>&1
ls -la gpurun_out/*.ncu-rep
