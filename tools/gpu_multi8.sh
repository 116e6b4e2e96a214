#!/bin/bash
# one-box scaling check: the driver's launch line for N ranks
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nproc; nvidia-smi -L | wc -l
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/r1b_bench_${N}gpu_b1.json 2> gpurun_out/${N}gpu_b1.err; tail -1 gpurun_out/r1b_bench_${N}gpu_b1.json | cut -c1-200; grep -v "OMP_NUM\|\*\*\*" gpurun_out/${N}gpu_b1.err | tail -3
python - <<PY
import json
try:
    r=json.loads(open('gpurun_out/r1b_bench_${N}gpu_b1.json').read().strip().splitlines()[-1])
    print('N=$N value %.0f e2e %.0f'%(r['value'], r['e2e']['value']), r['clocks'])
except Exception as e: print('ERR', e)
PY
