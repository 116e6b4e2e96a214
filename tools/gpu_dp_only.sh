#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29510 tools/dp_check.py > gpurun_out/r2_dp_check_${N}gpu.json 2> gpurun_out/dp_check_${N}gpu.err; tail -1 gpurun_out/r2_dp_check_${N}gpu.json | cut -c1-1100; tail -2 gpurun_out/dp_check_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload train --steps 20 --warmup 5 > gpurun_out/r2_bench_train_${N}gpu.json 2> gpurun_out/bench_train_${N}gpu.err; python - <<PY
import json
v=json.loads(open('gpurun_out/r2_bench_train_${N}gpu.json').read().strip().splitlines()[-1])
print('train', round(v['value']), v['unit'], 'e2e', round(v['e2e']['value']), 'ms/step', v['ms_per_step'], json.dumps(v.get('collective'))[300:])
PY
