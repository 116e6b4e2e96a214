#!/bin/bash
mkdir -p gpurun_out
for cfg in "conv3_2+pool 256,12" "conv3_2+pool 256,11" "conv2_2+pool 128,12" "conv2_2+pool 128,11" "conv4_2+pool 192,12"; do
  set -- $cfg
  echo "== $1 cfg $2 batch ${B:-1}"
  FRCNN_CONV_TRACE=gpurun_out/conv_trace.bin FRCNN_BENCH_LAYER=$1 FRCNN_BENCH_CFG="$2" python tools/bench_conv_layers.py ${B:-1} 2>&1 | tail -1
  python tools/conv_trace.py gpurun_out/conv_trace.bin 2>/dev/null | head -6
done
