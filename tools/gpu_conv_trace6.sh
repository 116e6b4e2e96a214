#!/bin/bash
mkdir -p gpurun_out
for dbg in 26 42 58; do
  echo "== dbg=$dbg conv2_1 128,51 batch 8"
  FRCNN_CONV_DBG=$dbg FRCNN_CONV_TRACE=gpurun_out/conv_trace.bin FRCNN_BENCH_LAYER=conv2_1 FRCNN_BENCH_CFG="128,51" python tools/bench_conv_layers.py 8 2>&1 | tail -1
  python tools/conv_trace.py gpurun_out/conv_trace.bin 2>/dev/null | sed -n 3,5p
done
cuobjdump -sass -fun '_ZN5frcnn21conv_pair_bres_kernelILi128EEEvNS_8ConvMapsENS_9ConvGroupE' faster-rcnn.torch_b200/build/conv_igemm.o | grep -B30 -A10 "UTCHMMA" | head -150 > gpurun_out/bres_sass.txt
