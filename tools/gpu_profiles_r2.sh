#!/bin/bash
# Round-2 evidence on one GPU: the default bench line (detect headline + nms / train / vgg_large nested), the reference arm,
# ncu launch lists with SM-active time (throughput schedule, batch 1 / batch 8 / vgg_large / latency schedule), one ncu
# --set full capture of the tcgen05 kernels of a batch-1 step, stage costs, the NMS sweep and the next-row micro-benchmarks.
# Only small text files go back (gpurun_out is capped at 64 MiB).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=r2
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -c 200 gpurun_out/${T}_bench_default.json
timeout 300 python bench.py --no-cpu-baseline --workload detect --batch 8 --in-flight 3 --steps 60 > gpurun_out/${T}_bench_b8.json 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline --workload detect --in-flight 1 --steps 60 > gpurun_out/${T}_bench_b1_sync.json 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline --workload detect --schedule latency --steps 200 > gpurun_out/${T}_bench_b1_latency_sched.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2>/dev/null
M=gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics $M --clock-control none -c 300 --csv --log-file gpurun_out/smtime_${T}_b1.csv python bench.py --workload detect --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 > /dev/null 2>&1
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics $M --clock-control none -c 300 --csv --log-file gpurun_out/smtime_${T}_b8.csv python bench.py --workload detect --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 --batch 8 > /dev/null 2>&1
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics $M --clock-control none -c 300 --csv --log-file gpurun_out/smtime_${T}_b1_latency.csv python bench.py --workload detect --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 --schedule latency > /dev/null 2>&1
FRCNN_NO_GRAPH=1 timeout 600 ncu --metrics $M --clock-control none -c 300 --csv --log-file gpurun_out/smtime_${T}_large_b1.csv python bench.py --workload detect --model vgg_large --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 > /dev/null 2>&1
for n in b1 b8 b1_latency large_b1; do python tools/smtime.py gpurun_out/smtime_${T}_$n.csv > gpurun_out/${T}_smtime_$n.txt; tail -1 gpurun_out/${T}_smtime_$n.txt; done
# full capture of the tcgen05 launches of one batch-1 detect step (first layer, 6 trunk convs, anchor networks, 2 cnet GEMMs)
for b in 1 8; do
  FRCNN_NO_GRAPH=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"conv_halo|conv_igemm|conv_first|conv_pair|conv_head" -s 70 -c 10 -o /tmp/${T}_full_b$b -f python bench.py --workload detect --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1 --batch $b > /dev/null 2>&1
  ncu -i /tmp/${T}_full_b$b.ncu-rep --page raw --csv > gpurun_out/${T}_full_b$b.csv 2>/dev/null
  ls -la /tmp/${T}_full_b$b.ncu-rep gpurun_out/${T}_full_b$b.csv
done
bash tools/stage_costs.sh --in-flight 8 > /dev/null; cp gpurun_out/stage_costs.txt gpurun_out/${T}_stage_costs.txt; cat gpurun_out/${T}_stage_costs.txt
timeout 300 python tools/bench_next_rows.py > gpurun_out/${T}_next_rows.jsonl 2>&1; cut -c1-300 gpurun_out/${T}_next_rows.jsonl
du -sh gpurun_out
