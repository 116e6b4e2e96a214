#!/bin/bash
for smt in 1 0; do for k in 4 6 8 12 16; do
FRCNN_CONV_SMTIME=$smt python bench.py --workload detect --steps 200 --warmup 5 --no-cpu-baseline --in-flight $k 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('smtime=$smt in_flight=$k us_per_frame %.1f  img/s %.0f  e2e %.0f clocks %s' % (1e3*d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']))"
done; done
