"""Turns the files tools/gpu_profiles.sh brought back in gpurun_out/ into the tracked summaries under profiles/ (r1b_*)."""
import csv
import json
import os
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from parse_launches import parse  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = "r1b"


def launches(name, title, cmd):
    step = parse(os.path.join(G, "launches_%s_%s.csv" % (TAG, name)))
    tot = sum(r[2] for r in step)
    agg = {}
    for r in step:
        a = agg.setdefault(r[1], [0, 0.0])
        a[0] += 1
        a[1] += r[2]
    with open(os.path.join(P, "%s_launches_%s.md" % (TAG, name)), "w") as o:
        o.write("# ncu launch list: %s\n\n`%s` (cold-cache, serialised launches: compare SHARES).  %d launches, %.1f us in total.\n\n"
                % (title, cmd, len(step), tot))
        o.write("| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write("| %s | %d | %.1f | %.1f%% |\n" % (k, n, t, 100 * t / tot))
        o.write("\n## in launch order\n\n| # | kernel | us | grid |\n|---:|---|---:|---|\n")
        for i, r in enumerate(step):
            o.write("| %d | %s | %.1f | %s |\n" % (i, r[1], r[2], r[3]))
    return tot


LAYERS = ["conv1_1 + pool (first-layer kernel)", "conv2_1", "conv2_2 + pool", "conv3_1", "conv3_2 + pool", "conv4_1", "conv4_2 + pool",
          "4 anchor heads, fused (k x k + tail)", "cnet fc1 (split-K reduce)", "cnet fc2 (split-K reduce)"]
GFLOP1 = [1.244, 13.271, 26.542, 13.330, 26.660, 10.086, 15.129, 24.052 + 0.082, None, None]


def full(batch):
    rows = list(csv.reader(open(os.path.join(G, "%s_full_b%d.csv" % (TAG, batch)))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = lambda n: hdr.index(n)  # noqa: E731

    def val(r, n, scale_units):
        v, u = float(r[col(n)]), units[col(n)].lower()
        return v * scale_units.get(u, 1.0)

    out = []
    with open(os.path.join(P, "%s_ncu_full_b%d.md" % (TAG, batch)), "w") as o:
        o.write("# ncu --set full: the ten tcgen05 launches of one pipelined `Detector:detect` step (throughput schedule), vgg_small 800x450, batch %d\n\n" % batch)
        o.write("`FRCNN_NO_GRAPH=1 ncu --set full --clock-control none -k regex:\"conv_halo|conv_igemm|conv_first\" -s 70 -c 10 python bench.py --steps 2 "
                "--warmup 1 --no-cpu-baseline --in-flight 1%s` (exported with `ncu -i ... --page raw --csv`).  DRAM traffic = dram__bytes_read.sum + "
                "dram__bytes_write.sum per launch; algorithmic FLOPs = 2*Cin*Cout*k^2*Hout*Wout*N.  Kernel times under ncu are cold-cache and serialised; "
                "the bench line's time is what counts.\n\n" % ("" if batch == 1 else " --batch %d" % batch))
        o.write("| layer | kernel | grid | us | GFLOP | TFLOP/s | tensor pipe active % | issue active % | L2 hit % | dram read MB | dram write MB | regs |\n"
                "|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for name, gf, r in zip(LAYERS, GFLOP1, data):
            us = val(r, "gpu__time_duration.sum", {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3})
            rd = val(r, "dram__bytes_read.sum", {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3})
            wr = val(r, "dram__bytes_write.sum", {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3})
            kern = r[col("Kernel Name")].split("(")[0].replace("void ", "").replace("frcnn::", "")
            g = gf * batch if gf else None
            o.write("| %s | %s | %s | %.1f | %s | %s | %.1f | %.1f | %.1f | %.2f | %.2f | %s |\n"
                    % (name, kern, r[col("Grid Size")], us, "%.2f" % g if g else "-", "%.0f" % (g / us * 1e3) if g else "-",
                       float(r[col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")]),
                       float(r[col("smsp__issue_active.avg.pct_of_peak_sustained_active")]), float(r[col("lts__t_sector_hit_rate.pct")]),
                       rd, wr, r[col("launch__registers_per_thread")]))
            out.append(dict(layer=name, us=us, gflop=g, dram_mb=rd + wr,
                            tensor_pct=float(r[col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")])))
    return out


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    for f in os.listdir(G):
        if f.startswith(TAG + "_bench_") and f.endswith(".json"):
            txt = open(os.path.join(G, f)).read().strip().splitlines()
            if txt:
                open(os.path.join(P, f), "w").write(txt[-1] + "\n")
    if os.path.exists(os.path.join(G, TAG + "_next_rows.jsonl")):
        shutil.copy(os.path.join(G, TAG + "_next_rows.jsonl"), os.path.join(P, TAG + "_next_rows.jsonl"))
    base = "FRCNN_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1"
    summ = {}
    summ["launch_b1_us"] = launches("b1", "one pipelined `Detector:detect` step (throughput schedule), vgg_small 800x450, batch 1", base)
    summ["launch_b8_us"] = launches("b8", "one pipelined `Detector:detect` step (throughput schedule), vgg_small 800x450, batch 8", base + " --batch 8")
    summ["launch_b1_latency_us"] = launches("b1_latency", "one `Detector:detect` step on the latency schedule, vgg_small 800x450, batch 1",
                                            base + " --schedule latency")
    summ["full_b1"] = full(1)
    summ["full_b8"] = full(8)
    json.dump(summ, open(os.path.join(P, TAG + "_summary.json"), "w"), indent=1)
    print("ok")
