#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r1b_bench_2gpu_b1.json 2> gpurun_out/2gpu_b1.err; tail -1 gpurun_out/r1b_bench_2gpu_b1.json | cut -c1-300; tail -3 gpurun_out/2gpu_b1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --batch 8 --steps 5 --warmup 3 > gpurun_out/r1b_bench_2gpu_train.json 2> gpurun_out/2gpu_train.err; tail -1 gpurun_out/r1b_bench_2gpu_train.json | cut -c1-300; tail -3 gpurun_out/2gpu_train.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload nms --steps 10 --warmup 3 > gpurun_out/r1b_bench_2gpu_nms.json 2> gpurun_out/2gpu_nms.err; tail -1 gpurun_out/r1b_bench_2gpu_nms.json | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
