#!/bin/bash
# power / clock trace (20 ms period) across a detect bench run with 8 frames in flight, then the stage-cost table
mkdir -p gpurun_out
nvidia-smi --query-gpu=timestamp,power.draw,clocks.sm,clocks.mem,temperature.gpu,clocks_throttle_reasons.sw_power_cap,clocks_throttle_reasons.hw_slowdown,power.limit --format=csv -lms 20 > gpurun_out/power_trace.csv 2>&1 &
SMI=$!
sleep 1
python bench.py --workload detect --steps 2000 --warmup 5 --no-cpu-baseline --in-flight 8 --min-seconds 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('detect us_per_frame %.1f  img/s %.0f' % (1e3*d['ms_per_step'], d['value']))"
sleep 1
kill $SMI
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/power_trace.csv')))
hdr=rows[0]; data=[r for r in rows[1:] if len(r)==len(hdr)]
pw=[float(r[1].split()[0]) for r in data]; ck=[float(r[2].split()[0]) for r in data]
busy=[i for i,p in enumerate(pw) if p>400]
print('samples',len(data),'busy',len(busy),'power limit',data[0][7])
if busy:
    import statistics as st
    print('power busy: median %.0f W max %.0f W' % (st.median(pw[i] for i in busy), max(pw[i] for i in busy)))
    print('sm clock busy: median %.0f min %.0f' % (st.median(ck[i] for i in busy), min(ck[i] for i in busy)))
    print('sw_power_cap active share', sum(1 for i in busy if 'Active' in data[i][5] and 'Not' not in data[i][5])/len(busy))
PY
bash tools/stage_costs.sh --in-flight 8
