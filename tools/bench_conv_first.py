"""Device time of the first-layer kernel alone (frcnn_conv_first, `iters` back-to-back launches) at 800x450.
FRCNN_FIRST_TMA=0 selects the register-gather kernel."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import frcnn_b200 as F  # noqa: E402

m = F.vgg_small(F.duplo_cfg)
ffi, L = F.ffi, F.lib()
for n in [int(a) for a in sys.argv[1:]] or [1, 8]:
    x = torch.randn(n, 3, 450, 800, device="cuda")
    w = torch.randn(64, 3, 3, 3, device="cuda") * 0.2
    b = torch.randn(64, device="cuda") * 0.1
    s = torch.tensor([0.25], device="cuda")
    out = torch.empty(n, 225, 400, 64, dtype=torch.bfloat16, device="cuda")
    ms = ffi.new("float*")
    best = None
    for _ in range(3):
        rc = L.frcnn_conv_first(m.ctx, ffi.cast("const float*", x.data_ptr()), ffi.cast("const float*", w.data_ptr()),
                                ffi.cast("const float*", b.data_ptr()), ffi.cast("const float*", s.data_ptr()), 1.0, n, 450, 800, 64, 1, 1,
                                ffi.cast("uint16_t*", out.data_ptr()), 20, ms)
        assert rc == 0, ffi.string(L.frcnn_last_error(m.ctx))
        us = ms[0] * 1000 / 20
        best = us if best is None else min(best, us)
    print(json.dumps({"tma": os.environ.get("FRCNN_FIRST_TMA", "1"), "batch": n, "us": round(best, 2)}))
m.close()
