"""Per-layer timing of the trunk convolutions (vgg_small at 800x450) through frcnn_conv_bf16: device time of `iters`
back-to-back launches of the conv kernel alone.  Run once per kernel variant:
    FRCNN_CONV_HALO=0 python tools/bench_conv_layers.py     # tap-per-box kernel
    python tools/bench_conv_layers.py                        # halo-tile kernel where eligible
Optional argv: batch sizes (default 1 8)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import frcnn_b200 as F  # noqa: E402

LAYERS = [  # name, h, w, cin, cout, pool
    ("conv2_1", 225, 400, 64, 128, 0), ("conv2_2+pool", 225, 400, 128, 128, 1), ("conv3_1", 113, 200, 128, 256, 0),
    ("conv3_2+pool", 113, 200, 256, 256, 1), ("conv4_1", 57, 100, 256, 384, 0), ("conv4_2+pool", 57, 100, 384, 384, 1),
]


LARGE = [  # vgg_large at 1000x600: the layers vgg_small does not have (FRCNN_BENCH_MODEL=large)
    ("conv1_2+pool", 600, 1000, 64, 64, 1), ("conv2_2+pool", 300, 500, 128, 128, 1), ("conv3_3+pool", 150, 250, 256, 256, 1),
    ("conv4_1", 75, 125, 256, 512, 0), ("conv4_3+pool", 75, 125, 512, 512, 1),
]


def main():
    global LAYERS
    if os.environ.get("FRCNN_BENCH_MODEL") == "large":
        LAYERS = LARGE
    batches = [int(a) for a in sys.argv[1:]] or [1, 8]
    m = F.vgg_small(F.duplo_cfg)
    ffi, L = F.ffi, F.lib()
    rows = []
    for n in batches:
        for name, h, w, cin, cout, pool in LAYERS:
            if os.environ.get("FRCNN_BENCH_LAYER") and os.environ["FRCNN_BENCH_LAYER"] != name:
                continue
            x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
            wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
            b = torch.zeros(cout, device="cuda")
            s = torch.tensor([0.25], device="cuda")
            ho, wo = ((h + 1) // 2, (w + 1) // 2) if pool else (h, w)
            out = torch.empty(n, ho, wo, cout, dtype=torch.bfloat16, device="cuda")
            ms = ffi.new("float*")
            cfgs = [tuple(int(v) for v in c.split(",")) for c in os.environ["FRCNN_BENCH_CFG"].split(";")] if os.environ.get("FRCNN_BENCH_CFG") else [(0, 0)]
            for forced in cfgs:
                if forced[0] and cout % forced[0]:
                    continue
                best = None
                for rep in range(int(os.environ.get("FRCNN_BENCH_REPS", "3"))):
                    rc = L.frcnn_conv_bf16(m.ctx, ffi.cast("const uint16_t*", x.data_ptr()), ffi.cast("const float*", wt.data_ptr()),
                                           ffi.cast("const float*", b.data_ptr()), ffi.cast("const float*", s.data_ptr()), 1.0, n, h, w,
                                           cin, cout, 3, 1, 0, forced[0], forced[1], pool, ffi.cast("uint16_t*", out.data_ptr()),
                                           int(os.environ.get("FRCNN_BENCH_ITERS", "20")), ms)
                    if rc != 0:
                        break
                    us = ms[0] * 1000.0 / int(os.environ.get("FRCNN_BENCH_ITERS", "20"))
                    best = us if best is None else min(best, us)
                if best is None:
                    continue
                gflop = 2.0 * n * h * w * cin * cout * 9 / 1e9
                rows.append({"batch": n, "layer": name, "cfg": list(forced), "us": round(best, 2), "tflops": round(gflop / best * 1e3, 1)})
                print(json.dumps(rows[-1]), flush=True)
    m.close()


if __name__ == "__main__":
    main()
