#!/bin/bash
# N-GPU evidence (N = $1): the C-ABI data-parallel check, then the driver's launch line of the default bench (detect
# headline + nms / train / vgg_large nested) and the training line alone
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29510 tools/dp_check.py > gpurun_out/r2_dp_check_${N}gpu.json 2> gpurun_out/dp_check_${N}gpu.err; tail -1 gpurun_out/r2_dp_check_${N}gpu.json | cut -c1-900; tail -2 gpurun_out/dp_check_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -1 gpurun_out/r2_bench_${N}gpu.json | cut -c1-300; grep -c "NCCL INFO" gpurun_out/bench_${N}gpu.err; grep -m3 -E "NVLS|nranks|Connected all" gpurun_out/bench_${N}gpu.err | cut -c1-200
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('detect', round(d['value']), 'e2e', round(d['e2e']['value']))
for k,v in d.get('also',{}).items():
    print(k, round(v['value']), v['unit'], 'e2e', round(v['e2e']['value']), json.dumps(v.get('collective'))[:600])
PY
