"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/ (round tag argv[1])."""
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from parse_launches import parse  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def launches(name, batch):
    step = parse(os.path.join(G, "launches_%s.csv" % name))
    tot = sum(r[2] for r in step)
    agg = {}
    for r in step:
        k = r[1]
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += r[2]
    with open(os.path.join(P, "%s_launches_%s.md" % (tag, name)), "w") as o:
        o.write("# ncu launch list: one `Detector:detect` step, vgg_small 800x450, batch %d\n\n" % batch)
        o.write("`FRCNN_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 1%s` "
                "(cold-cache, serialised launches: compare SHARES).  %d launches, %.1f us in total.\n\n"
                % ("" if batch == 1 else " --batch %d" % batch, len(step), tot))
        o.write("| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write("| %s | %d | %.1f | %.1f%% |\n" % (k, n, t, 100 * t / tot))
        o.write("\n## in launch order\n\n| # | kernel | us | grid |\n|---:|---|---:|---|\n")
        for i, r in enumerate(step):
            o.write("| %d | %s | %.1f | %s |\n" % (i, r[1], r[2], r[3]))
    conv = sum(t for k, (n, t) in agg.items() if k.startswith("conv_"))
    return tot, conv


def full(name, batch, layer_names, flops):
    rep = os.path.join(G, "prof_conv_%s.ncu-rep" % name)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[0], rows[2:]
    data = [r for r in data if "conv_igemm_kernel" in r[hdr.index("Kernel Name")] or "conv_first_kernel" in r[hdr.index("Kernel Name")]]
    first = [i for i, r in enumerate(data) if "conv_first_kernel" in r[hdr.index("Kernel Name")]][0]
    data = data[first:]  # one step starts with the first-layer kernel
    want = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__cycles_elapsed.avg"]
    idx = [hdr.index(w) for w in want]
    units = rows[1]
    out = []
    with open(os.path.join(P, "%s_ncu_conv_%s.md" % (tag, name)), "w") as o:
        o.write("# ncu --set full: the conv kernels of one `Detector:detect` step, vgg_small 800x450, batch %d\n\n" % batch)
        o.write("`FRCNN_NO_GRAPH=1 ncu --set full --clock-control none -k regex:conv_igemm|conv_first -s 9 -c 9 python bench.py --steps 1 --warmup 1%s`; "
                "per-launch DRAM traffic = dram__bytes_read.sum + dram__bytes_write.sum; algorithmic FLOPs = 2*Cin*Cout*k^2*Hout*Wout*N.\n\n"
                % ("" if batch == 1 else " --batch %d" % batch))
        o.write("| layer | kernel | grid | us | GFLOP | TFLOP/s | tensor pipe %% (active) | dram read MB | dram write MB | regs |\n"
                "|---|---|---|---:|---:|---:|---:|---:|---:|---:|\n")
        for n, fl, r in zip(layer_names, flops, data):
            v = [r[i] for i in idx]
            us = float(v[2]) * (1e-3 if units[idx[2]] in ("ns", "nsecond") else 1.0) * (1e3 if units[idx[2]] in ("ms", "msecond") else 1.0)
            kern = v[0].split("(")[0].replace("void ", "").replace("frcnn::", "").replace("(int)", "")
            rd, wr = float(v[3]), float(v[4])
            if units[idx[3]].lower().startswith("kbyte"): rd /= 1e3
            if units[idx[4]].lower().startswith("kbyte"): wr /= 1e3
            if units[idx[3]].lower() == "byte": rd /= 1e6
            if units[idx[4]].lower() == "byte": wr /= 1e6
            o.write("| %s | %s | %s | %.1f | %.2f | %.0f | %.1f | %.2f | %.2f | %s |\n"
                    % (n, kern, v[1], us, fl / 1e9, fl / us / 1e6, float(v[5]), rd, wr, v[8]))
            out.append(dict(layer=n, us=us, gflop=fl / 1e9, dram_mb=rd + wr, tensor_pct=float(v[5])))
    return out


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    names = ["conv1_1 (fused first layer + pool)", "conv2_1", "conv2_2 + pool", "conv3_1", "conv3_2 + pool", "conv4_1", "conv4_2 + pool",
             "4 anchor heads (grouped, split-K slices)", "fc1 (split-K reduce)"]
    per_img = [1.244e9, 13.271e9, 26.542e9, 13.330e9, 26.660e9, 10.086e9, 15.129e9, 6.358e9 + 2.293e9 + 5.652e9 + 9.749e9]
    res = {}
    for name, batch in (("b1", 1), ("b8", 8)):
        if os.path.exists(os.path.join(G, "launches_%s.csv" % name)):
            res["launch_%s" % name] = launches(name, batch)
        if os.path.exists(os.path.join(G, "prof_conv_%s.ncu-rep" % name)):
            fl = [f * batch for f in per_img] + [0.0]
            res["full_%s" % name] = full(name, batch, names, fl)
    json.dump(res, open(os.path.join(P, "%s_summary.json" % tag), "w"), indent=1)
    print(json.dumps(res, indent=1)[:3000])
