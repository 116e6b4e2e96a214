#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 2; do
for cfg in "conv2_1 128,51" "conv2_1 128,21" "conv2_2+pool 128,51"; do
  set -- $cfg
  echo "== dbg=$dbg $1 cfg $2 batch 8"
  FRCNN_CONV_DBG=$dbg FRCNN_CONV_TRACE=gpurun_out/conv_trace.bin FRCNN_BENCH_LAYER=$1 FRCNN_BENCH_CFG="$2" python tools/bench_conv_layers.py 8 2>&1 | tail -1
  python tools/conv_trace.py gpurun_out/conv_trace.bin 2>/dev/null | sed -n 2,6p
done; done
