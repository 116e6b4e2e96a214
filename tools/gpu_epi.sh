#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_conv.py tests/test_gpu_detect.py tests/test_gpu_train.py tests/test_gpu_conv_backward.py -m gpu -q -x 2>&1 | tail -4
for cfg in "conv2_1 128,31" "conv2_2+pool 128,31" "conv3_2+pool 256,12"; do
  set -- $cfg
  for dbg in 0 64; do
  echo "== dbg=$dbg $1 cfg $2 batch 1"
  FRCNN_CONV_DBG=$dbg FRCNN_CONV_TRACE=gpurun_out/conv_trace.bin FRCNN_BENCH_LAYER=$1 FRCNN_BENCH_CFG="$2" python tools/bench_conv_layers.py 1 8 2>&1 | tail -2
  python tools/conv_trace.py gpurun_out/conv_trace.bin 2>/dev/null | sed -n 3,4p
  done
done
python bench.py --workload detect --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('detect us_per_frame %.1f  img/s %.0f  e2e %.0f  sync_us %.1f frac %.3f' % (1e3*d['ms_per_step'], d['value'], d['e2e']['value'], 1e3*d['config']['sync']['ms_per_step'], d['roofline']['frac']))"
python bench.py --workload train --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train', round(d['value']), 'e2e', round(d['e2e']['value']))"
