#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max --clock-control none --csv --log-file gpurun_out/sweep_smtime.csv -k regex:'conv_(halo|igemm|pair|pair_bres)_kernel' python tools/sweep_smtime.py run "$@" > gpurun_out/sweep_smtime.log 2>&1
python tools/sweep_smtime.py join gpurun_out/sweep_smtime.csv gpurun_out/sweep_smtime_order.jsonl | tee gpurun_out/sweep_smtime.md
