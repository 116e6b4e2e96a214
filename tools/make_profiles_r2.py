"""Turns the files tools/gpu_profiles_r2.sh brought back in gpurun_out/ into the tracked summaries under profiles/ (r2_*)."""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = "r2"

GFLOP1 = {"conv_first": 1.244, "conv2_1": 13.271, "conv2_2": 26.542, "conv3_1": 13.330, "conv3_2": 26.660, "conv4_1": 10.086,
          "conv4_2": 15.129, "heads": 24.052}
ORDER = ["conv_first", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv4_1", "conv4_2", "heads", "fc1", "fc2"]
NAMES = {"conv_first": "conv1_1 + pool (first-layer kernel)", "conv2_1": "conv2_1", "conv2_2": "conv2_2 + pool", "conv3_1": "conv3_1",
         "conv3_2": "conv3_2 + pool", "conv4_1": "conv4_1", "conv4_2": "conv4_2 + pool", "heads": "4 anchor networks (k x k, pairs)",
         "fc1": "cnet fc1 (split-K slices)", "fc2": "cnet fc2 (split-K slices)"}


def full(batch):
    path = os.path.join(G, "%s_full_b%d.csv" % (TAG, batch))
    if not os.path.exists(path):
        return []
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = lambda n: hdr.index(n)  # noqa: E731

    def val(r, n, scale_units):
        v, u = float(r[col(n)].replace(",", "")), units[col(n)].lower()
        return v * scale_units.get(u, 1.0)

    out = []
    with open(os.path.join(P, "%s_ncu_full_b%d.md" % (TAG, batch)), "w") as o:
        o.write("# ncu --set full: the ten tcgen05 launches of one `Detector:detect` step (throughput schedule), vgg_small 800x450, batch %d (round 2)\n\n" % batch)
        o.write("`FRCNN_NO_GRAPH=1 ncu --set full --import-source on --clock-control none -k regex:\"conv_halo|conv_igemm|conv_first|conv_pair|conv_head\" -s 70 -c 10 "
                "python bench.py --workload detect --steps 2 --warmup 1 --no-cpu-baseline --in-flight 1%s` (exported with `ncu -i ... --page raw --csv`).  DRAM traffic = "
                "dram__bytes_read.sum + dram__bytes_write.sum per launch; algorithmic FLOPs = 2*Cin*Cout*k^2*Hout*Wout*N.  Kernel times under ncu are cold-cache and "
                "serialised; `SM-active us` = sm__cycles_active.avg / clock = mean CTA residency, what the launch costs a frame when other frames fill the idle SMs.\n\n"
                % ("" if batch == 1 else " --batch %d" % batch))
        o.write("| layer | kernel | grid | us | SM-active us | GFLOP | TFLOP/s | TFLOP/s of SM-time | tensor pipe active % | issue active % | L2 hit % | dram read MB | dram write MB | regs |\n"
                "|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for key, r in zip(ORDER, data):
            us = val(r, "gpu__time_duration.sum", {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3})
            rd = val(r, "dram__bytes_read.sum", {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3})
            wr = val(r, "dram__bytes_write.sum", {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3})
            act = float(r[col("sm__cycles_active.avg")].replace(",", "")) / float(r[col("sm__cycles_elapsed.max")].replace(",", "")) * us
            kern = r[col("Kernel Name")].split("(")[0].replace("void ", "").replace("frcnn::", "")
            g = GFLOP1[key] * batch if key in GFLOP1 else None
            o.write("| %s | %s | %s | %.1f | %.1f | %s | %s | %s | %.1f | %.1f | %.1f | %.2f | %.2f | %s |\n"
                    % (NAMES[key], kern, r[col("Grid Size")], us, act, "%.2f" % g if g else "-", "%.0f" % (g / us * 1e3) if g else "-",
                       "%.0f" % (g / act * 1e3) if g else "-",
                       float(r[col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")]),
                       float(r[col("smsp__issue_active.avg.pct_of_peak_sustained_active")]), float(r[col("lts__t_sector_hit_rate.pct")]),
                       rd, wr, r[col("launch__registers_per_thread")]))
            out.append(dict(layer=NAMES[key], us=us, sm_active_us=act, gflop=g, dram_mb=rd + wr))
    return out


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    for f in os.listdir(G):
        if f.startswith(TAG + "_bench_") and f.endswith(".json"):
            txt = open(os.path.join(G, f)).read().strip().splitlines()
            if txt:
                open(os.path.join(P, f), "w").write(txt[-1] + "\n")
    for f in ("next_rows.jsonl", "stage_costs.txt", "smtime_b1.txt", "smtime_b8.txt", "smtime_b1_latency.txt", "smtime_large_b1.txt"):
        src = os.path.join(G, "%s_%s" % (TAG, f))
        if os.path.exists(src):
            shutil.copy(src, os.path.join(P, "%s_%s" % (TAG, f)))
    summ = {"full_vgg_small_b1": full(1), "full_vgg_small_b8": full(8)}   # the keys bench.py looks up for roofline.traffic
    json.dump(summ, open(os.path.join(P, TAG + "_summary.json"), "w"), indent=1)
    print("ok")
