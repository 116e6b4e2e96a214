#!/bin/bash
# precision tests + ncu launch list of one synchronous batch-1 step (throughput schedule, the bench's contexts)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_precision.py -m gpu -q -x > gpurun_out/pytest_prec.log 2>&1; echo "pytest rc=$?"; grep -h '"frames"' gpurun_out/pytest_prec.log | cut -c1-700; tail -3 gpurun_out/pytest_prec.log
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2_b1.csv python bench.py --workload detect --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 > /dev/null 2>&1
python tools/parse_launches.py gpurun_out/launches_r2_b1.csv
