"""SM-time sweep of the trunk convolutions: which kernel / tile configuration costs the fewest SM-microseconds per launch
(what a frame costs when the SMs a launch leaves idle are filled by the other frames in flight), next to the isolated
duration.  Two passes on the GPU box:
    ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max --clock-control none --csv \
        --log-file gpurun_out/sweep_smtime.csv -k regex:'conv_(halo|igemm|pair|pair_bres)_kernel' python tools/sweep_smtime.py run [batch]
    python tools/sweep_smtime.py join gpurun_out/sweep_smtime.csv gpurun_out/sweep_smtime_order.jsonl
`run` launches every configuration exactly once (in the order it logs), `join` pairs the ncu rows with that log."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LAYERS = [("conv2_1", 225, 400, 64, 128, 0), ("conv2_2+pool", 225, 400, 128, 128, 1), ("conv3_1", 113, 200, 128, 256, 0),
          ("conv3_2+pool", 113, 200, 256, 256, 1), ("conv4_1", 57, 100, 256, 384, 0), ("conv4_2+pool", 57, 100, 384, 384, 1)]
LARGE = [("conv1_2+pool", 600, 1000, 64, 64, 1), ("conv2_2+pool", 300, 500, 128, 128, 1), ("conv3_3+pool", 150, 250, 256, 256, 1),
         ("conv4_1", 75, 125, 256, 512, 0), ("conv4_3+pool", 75, 125, 512, 512, 1)]
KINDS = {1: "tap x1", 2: "tap x2", 11: "halo x1", 12: "halo x2", 21: "pair x1", 22: "pair x2", 31: "halo x1 occ2", 32: "halo x2 occ2",
         41: "pair x1 occ2", 42: "pair x2 occ2", 51: "pair resident-B", 61: "swap 128x256"}


def run():
    import torch
    import frcnn_b200 as F
    batches = [int(a) for a in sys.argv[2:]] or [1]
    layers = LARGE if os.environ.get("FRCNN_BENCH_MODEL") == "large" else LAYERS
    m = F.vgg_small(F.duplo_cfg)
    ffi, L = F.ffi, F.lib()
    out_log = open(os.path.join(ROOT, "gpurun_out", "sweep_smtime_order.jsonl"), "w")
    for n in batches:
        for name, h, w, cin, cout, pool in layers:
            x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
            wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
            b = torch.zeros(cout, device="cuda")
            s = torch.tensor([0.25], device="cuda")
            ho, wo = ((h + 1) // 2, (w + 1) // 2) if pool else (h, w)
            out = torch.empty(n, ho, wo, cout, dtype=torch.bfloat16, device="cuda")
            ms = ffi.new("float*")
            for bn in (64, 128, 192, 256):
                if cout % bn:
                    continue
                for kind in sorted(KINDS):
                    rc = L.frcnn_conv_bf16(m.ctx, ffi.cast("const uint16_t*", x.data_ptr()), ffi.cast("const float*", wt.data_ptr()),
                                           ffi.cast("const float*", b.data_ptr()), ffi.cast("const float*", s.data_ptr()), 1.0, n, h, w,
                                           cin, cout, 3, 1, 0, bn, kind, pool, ffi.cast("uint16_t*", out.data_ptr()), 1, ms)
                    torch.cuda.synchronize()
                    if rc == 0:
                        out_log.write(json.dumps(dict(batch=n, layer=name, bn=bn, kind=kind, gflop=2.0 * n * h * w * cin * cout * 9 / 1e9)) + "\n")
                        out_log.flush()
    m.close()


def join():
    lines = [l for l in open(sys.argv[2]) if not l.startswith("==")]
    rows = {}
    for r in csv.DictReader(lines):
        d = rows.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]), "grid": r["Grid Size"]})
        d[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    launches = [rows[k] for k in sorted(rows)]
    order = [json.loads(l) for l in open(sys.argv[3])]
    assert len(launches) == len(order), (len(launches), len(order))
    best = {}
    for cfg, d in zip(order, launches):
        dur, u = d["gpu__time_duration.sum"]
        dur = dur / 1e3 if u in ("ns", "nsecond") else dur
        act = d["sm__cycles_active.avg"][0] / d["sm__cycles_elapsed.max"][0] * dur
        cfg.update(us=round(dur, 1), sm_us=round(act, 1), grid=d["grid"], kernel=d["name"].split("<")[0].replace("void frcnn::", ""))
        best.setdefault((cfg["batch"], cfg["layer"]), []).append(cfg)
    print("| batch | layer | fewest SM-us | shortest alone | all (kind BN: us alone / SM-us) |\n|---:|---|---|---|---|")
    for (bt, layer), cfgs in best.items():
        a = min(cfgs, key=lambda c: c["sm_us"])
        b = min(cfgs, key=lambda c: c["us"])
        f = lambda c: "%s %d: %.1f / %.1f" % (KINDS[c["kind"]], c["bn"], c["us"], c["sm_us"])  # noqa: E731
        print("| %d | %s | %s (%.0f TFLOP/s of SM-time) | %s | %s |" % (bt, layer, f(a), a["gflop"] / a["sm_us"] * 1e3, f(b),
                                                                       "; ".join(f(c) for c in sorted(cfgs, key=lambda c: c["sm_us"]))))
    with open(os.path.join(ROOT, "gpurun_out", "sweep_smtime.jsonl"), "w") as o:
        for cfgs in best.values():
            for c in cfgs:
                o.write(json.dumps(c) + "\n")


if __name__ == "__main__":
    run() if sys.argv[1] == "run" else join()
