#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for halo in 1 0; do for dbg in 0 1 2 4 6; do
  echo "== halo=$halo dbg=$dbg"
  FRCNN_CONV_HALO=$halo FRCNN_CONV_DBG=$dbg timeout 200 python tools/bench_conv_layers.py 1 8 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    if 'layer' in r: print(r['batch'], r['layer'], r['us'], end=' | ')
print()"
done; done 2>&1 | tee gpurun_out/dbg_matrix.log
