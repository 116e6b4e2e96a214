#!/bin/bash
mkdir -p gpurun_out
for cfg in "conv2_1 128,31" "conv2_2+pool 128,31" "conv2_2+pool 128,11" "conv3_2+pool 256,12"; do
  set -- $cfg
  echo "== $1 cfg $2 batch ${B:-1}"
  FRCNN_CONV_TRACE=gpurun_out/conv_trace.bin FRCNN_BENCH_LAYER=$1 FRCNN_BENCH_CFG="$2" python tools/bench_conv_layers.py ${B:-1} 2>&1 | tail -1
  python tools/conv_trace.py gpurun_out/conv_trace.bin
done
