#!/bin/bash
# Round-end evidence: bench lines (1 GPU), ncu launch lists (recipe: gpu__time_duration, --clock-control none) of the
# bench command on both schedules, and one ncu --set full capture of the tcgen05 kernels of a step.
# Only small text files go back (gpurun_out is capped at 64 MiB): the .ncu-rep files are exported to CSV and deleted.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if [ "${1:-all}" != "ncu" ]; then
timeout 600 python bench.py > gpurun_out/r1b_bench_b1.json 2> gpurun_out/r1b_bench_b1.err; tail -c 300 gpurun_out/r1b_bench_b1.json
timeout 300 python bench.py --no-cpu-baseline --batch 8 --in-flight 3 --steps 60 > gpurun_out/r1b_bench_b8.json 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline --in-flight 1 --steps 60 > gpurun_out/r1b_bench_b1_sync.json 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline --schedule latency --steps 200 > gpurun_out/r1b_bench_b1_latency_sched.json 2>/dev/null
timeout 300 python bench.py --workload nms > gpurun_out/r1b_bench_nms.json 2>/dev/null
timeout 300 python bench.py --workload train --batch 8 --steps 5 --warmup 3 > gpurun_out/r1b_bench_train_b8.json 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline --model vgg_large --batch 1 --steps 100 > gpurun_out/r1b_bench_large_b1.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1b_bench_reference_arm.json 2>/dev/null
fi
# launch lists (cold cache, serialised): throughput schedule = the bench's own contexts
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1b_b1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 > /dev/null 2>&1
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1b_b8.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 --batch 8 > /dev/null 2>&1
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1b_b1_latency.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 --schedule latency > /dev/null 2>&1
# full capture of the 10 tcgen05 launches of one detect step on a pipeline context (first layer, 6 trunk convs, fused
# heads, 2 cnet GEMMs): the step after the sync-timed ones (-s skips the earlier launches of the process)
for b in 1 8; do
  FRCNN_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none -k regex:"conv_halo|conv_igemm|conv_first" -s 70 -c 10 -o /tmp/r1b_full_b$b -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 --batch $b > /dev/null 2>&1
  ncu -i /tmp/r1b_full_b$b.ncu-rep --page raw --csv > gpurun_out/r1b_full_b$b.csv 2>/dev/null
  ls -la /tmp/r1b_full_b$b.ncu-rep gpurun_out/r1b_full_b$b.csv
done
du -sh gpurun_out
timeout 300 python tools/bench_next_rows.py > gpurun_out/r1b_next_rows.jsonl 2>&1; cat gpurun_out/r1b_next_rows.jsonl | cut -c1-400
bash tools/gpu_nms_sweep.sh > /dev/null 2>&1; wc -l gpurun_out/r1b_nms_sweep.jsonl
bash tools/gpu_trainll.sh 2>&1 | tail -32 > gpurun_out/r1b_train_launches.txt; head -3 gpurun_out/r1b_train_launches.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
