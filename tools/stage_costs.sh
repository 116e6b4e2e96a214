#!/bin/bash
# Effective cost of every pipeline stage with frames in flight: throughput of Detector:detect truncated after stage k
# (FRCNN_DETECT_STOP), differences between consecutive rows = the SM time the stage really costs per frame.
# Usage (on the GPU box): bash tools/stage_costs.sh [extra bench.py flags]  ->  gpurun_out/stage_costs.txt
mkdir -p gpurun_out
: > gpurun_out/stage_costs.txt
for s in 1 2 3 4 5 0; do
  FRCNN_DETECT_STOP=$s python bench.py --workload detect --steps 100 --warmup 5 --no-cpu-baseline --min-seconds 0.3 "$@" 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('stop=$s us_per_frame %.1f  img/s %.0f  sync_us %.1f' % (1e3*d['ms_per_step'], d['value'], 1e3*d['config']['sync']['ms_per_step']))" >> gpurun_out/stage_costs.txt
done
cat gpurun_out/stage_costs.txt
