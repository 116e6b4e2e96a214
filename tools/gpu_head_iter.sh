#!/bin/bash
# anchor-network kernel iteration: the tests that exercise it, the per-unit trace and the stage-cost rows around it
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_detect.py tests/test_gpu_precision.py -m gpu -q -x > gpurun_out/pytest_head.log 2>&1; echo "pytest rc=$?"; grep -h '"frames"' gpurun_out/pytest_head.log | cut -c1-420; tail -4 gpurun_out/pytest_head.log
FRCNN_NO_GRAPH=1 FRCNN_HEAD_TRACE=gpurun_out/head_trace.bin timeout 300 python bench.py --workload detect --steps 3 --warmup 1 --no-cpu-baseline --in-flight 1 > /dev/null 2>gpurun_out/head_trace.err
python tools/head_trace.py gpurun_out/head_trace.bin 10
for s in 1 2 0; do
  FRCNN_DETECT_STOP=$s python bench.py --workload detect --steps 100 --warmup 5 --no-cpu-baseline --min-seconds 0.3 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('stop=$s us_per_frame %.1f  img/s %.0f  sync_us %.1f' % (1e3*d['ms_per_step'], d['value'], 1e3*d['config']['sync']['ms_per_step']))"
done
