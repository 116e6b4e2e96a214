#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "resident" 2>&1 | tail -5
FRCNN_BENCH_LAYER=conv2_1 FRCNN_BENCH_CFG="128,51;128,31;128,41" python tools/bench_conv_layers.py 1 8 2>&1 | tail -6
FRCNN_BENCH_LAYER="conv2_2+pool" FRCNN_BENCH_CFG="128,51;128,31;128,41" python tools/bench_conv_layers.py 1 8 2>&1 | tail -6
bash tools/gpu_sweep_smtime.sh 1 2>/dev/null | grep -E "conv2_" | cut -c1-420
