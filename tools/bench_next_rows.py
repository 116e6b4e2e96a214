"""Measurement of the SURVEY 8f rows built this round, next to their CPU restatements (JSON lines):
  rmsprop_step   HBM-bound: 20 algorithmic bytes per parameter (12 read + 8 written; the divided gradient is written
                 back as well: +4) over the flat vgg_small buffer, CUDA events, L2 flushed between iterations
  find_positive  8 ground-truth boxes on an 800x450 frame, IoU evaluations per second
  sample_negative 128 negatives against 8 boxes."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import frcnn_b200 as F  # noqa: E402
from oracle import anchors as OA, model as OM, optim as OO  # noqa: E402  (CPU baselines only)
from oracle.rect import Rect  # noqa: E402

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
m = F.vgg_small(F.duplo_cfg)
ffi, L = F.ffi, F.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

# ---- rmsprop
n = m.weights.numel()
g = torch.randn_like(m.weights) * 1e-3
st = torch.zeros_like(m.weights)
w = m.weights.clone()
evs = []
for it in range(13):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    rc = L.frcnn_rmsprop_step(m.ctx, ffi.cast("float*", w.data_ptr()), ffi.cast("float*", g.data_ptr()), ffi.cast("float*", st.data_ptr()), n,
                              256.0, 1e-3, 0.99, 1e-8, 0.0)
    b.record()
    assert rc == 0
    torch.cuda.synchronize()
    if it >= 3:
        evs.append(a.elapsed_time(b))
ms = float(np.median(evs))
wc, gc, sc = w.cpu().numpy(), g.cpu().numpy(), {"m": st.cpu().numpy()}
t0 = time.perf_counter()
OO.rmsprop_step(wc, OO.gradient_div(gc, 256.0), sc, learningRate=1e-3)
cpu_s = time.perf_counter() - t0
gbs = 24.0 * n / (ms * 1e-3) / 1e9
print(json.dumps({"row": "rmsprop_step (gradient:div + optim.rmsprop, fused)", "params": n, "ms": round(ms, 4), "algorithmic_bytes": 24 * n,
                  "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(gbs / peaks["hbm_gbs"], 3)},
                  "cpu_baseline": {"ms": round(cpu_s * 1e3, 2), "kind": "port", "cores": 1, "sample": "numpy fp32 restatement, same buffer"}}))

# ---- frame normalisation
from oracle import preprocess as OP  # noqa: E402
frame = torch.rand(3, 450, 800)
fd = frame.clone().cuda()
evs = []
for it in range(13):
    fd.copy_(frame)
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    m.normalize_frame(fd, rgb2yuv=True)
    b.record()
    torch.cuda.synchronize()
    if it >= 3:
        evs.append(a.elapsed_time(b))
ms = float(np.median(evs))
torch.set_num_threads(os.cpu_count() or 1)
t0 = time.perf_counter()
OP.normalize_frame(frame, rgb_to_yuv=True)
cpu_s = time.perf_counter() - t0
fb = frame.numel() * 4
print(json.dumps({"row": "normalize_frame (rgb2yuv + centering + scaling + contrastive 7), 800x450", "ms": round(ms, 4), "launches": 11,
                  "algorithmic_bytes": int(fb * 2 + fb * 2 * 2 / 3), "note": "call incl. its final stream synchronisation; latency-bound at this size (11 small launches)",
                  "roofline": {"bound": "hbm", "achieved": round((fb * 2 + fb * 4 / 3) / (ms * 1e-3) / 1e9, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                               "frac": round((fb * 2 + fb * 4 / 3) / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)},
                  "cpu_baseline": {"ms": round(cpu_s * 1e3, 2), "kind": "port", "cores": os.cpu_count(), "sample": "PyTorch-CPU restatement, same frame"}}))

# ---- labelling
rng = np.random.default_rng(0)
rois = []
for _ in range(8):
    bw, bh = rng.uniform(40, 300), rng.uniform(40, 250)
    x, y = rng.uniform(0, 800 - bw), rng.uniform(0, 450 - bh)
    rois.append({"rect": Rect(x, y, x + bw, y + bh)})
ga = F.Anchors(m)
oa = OA.Anchors(OM.VGG_SMALL["layers"], OM.VGG_SMALL["anchor_nets"], OM.CFG_DUPLO["scales"])
img = Rect(0, 0, 800, 450)
cand = sum((r["ux"] - r["lx"]) * (r["uy"] - r["ly"]) for roi in rois for r in oa.findRangesXY(roi["rect"], img))
flat = np.ascontiguousarray([list(r["rect"].unpack()) for r in rois], dtype=np.float64).reshape(-1)
clip = ffi.new("double[4]", [0, 0, 800, 450])
cap = 1 << 16
out, out_roi, n_out = ffi.new("frcnn_anchor_ref[]", cap), ffi.new("int[]", cap), ffi.new("int*")
ts = []
for it in range(23):
    t0 = time.perf_counter()
    rc = L.frcnn_find_positive(m.ctx, ffi.cast("const double*", flat.ctypes.data), 8, clip, 0.6, 0.3, 1, out, out_roi, cap, n_out)
    assert rc == 0
    if it >= 3:
        ts.append(time.perf_counter() - t0)
gpu_s = float(np.median(ts))
t0 = time.perf_counter()
want = oa.findPositive(rois, img, 0.6, 0.3, True)
cpu_s = time.perf_counter() - t0
print(json.dumps({"row": "find_positive (8 boxes, 800x450)", "iou_evaluations": int(cand), "matches": int(n_out[0]), "call_ms": round(gpu_s * 1e3, 4),
                  "iou_per_s": round(cand / gpu_s), "note": "host call incl. H2D of the boxes, launch, D2H of the match list (latency-bound: one CTA per box)",
                  "cpu_baseline": {"ms": round(cpu_s * 1e3, 1), "kind": "port", "cores": 1, "sample": "oracle restatement of the Lua loops, same boxes", "matches": len(want)}}))
rnd = rng.integers(0, 2 ** 32, 3 * 1000, dtype=np.uint64).astype(np.uint32)
ts = []
for it in range(23):
    t0 = time.perf_counter()
    got, used, fin = ga.sampleNegative(img, rois, 0.3, 128, rnd)
    if it >= 3:
        ts.append(time.perf_counter() - t0)
gpu_s = float(np.median(ts))
t0 = time.perf_counter()
want, trials = oa.sampleNegative(img, rois, 0.3, 128, rnd)
cpu_s = time.perf_counter() - t0
print(json.dumps({"row": "sample_negative (128 of 8 boxes)", "trials": int(used), "call_ms": round(gpu_s * 1e3, 4),
                  "cpu_baseline": {"ms": round(cpu_s * 1e3, 1), "kind": "port", "cores": 1, "sample": "oracle restatement, same random stream"}}))

# ---- nearby-aversion candidates (Anchors:findNearby per positive anchor, BatchIterator.lua:206-217)
positive = oa.findPositive(rois, img, 0.6, 0.3, True)
ts = []
for it in range(23):
    t0 = time.perf_counter()
    got = ga.nearbyNegative(positive, 0.3)
    if it >= 3:
        ts.append(time.perf_counter() - t0)
gpu_s = float(np.median(ts))
t0 = time.perf_counter()
want = []
for pz in positive:
    cx, cy = pz[0].center()
    want += [a for a in oa.findNearby(cx, cy) if Rect.IoU(pz[0], a) < 0.3]
cpu_s = time.perf_counter() - t0
print(json.dumps({"row": "find_nearby_negative (%d positives)" % len(positive), "entries": len(got), "call_ms": round(gpu_s * 1e3, 4),
                  "note": "host call incl. marshalling of the positives and building the returned anchor rects in Python",
                  "cpu_baseline": {"ms": round(cpu_s * 1e3, 1), "kind": "port", "cores": 1, "sample": "oracle restatement of the Lua loops", "entries": len(want)}}))

# ---- resize (image.scale 'bilinear', BatchIterator.lua:49-52): 1280x720 -> 800x450
src = torch.rand(3, 720, 1280)
sd = src.cuda()
evs = []
for it in range(13):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    o = m.scale_frame(sd, 800, 450)
    b.record()
    torch.cuda.synchronize()
    if it >= 3:
        evs.append(a.elapsed_time(b))
ms = float(np.median(evs))
t0 = time.perf_counter()
OP.scale_image(src.numpy(), 800, 450)
cpu_s = time.perf_counter() - t0
ab = src.numel() * 4 + 3 * 450 * 800 * 4 + 2 * 3 * 720 * 800 * 4     # source read, result written, row-pass temporary written + read
print(json.dumps({"row": "scale_frame (image.scale bilinear) 1280x720 -> 800x450", "ms": round(ms, 4), "launches": 2, "algorithmic_bytes": int(ab),
                  "roofline": {"bound": "hbm", "achieved": round(ab / (ms * 1e-3) / 1e9, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                               "frac": round(ab / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)},
                  "cpu_baseline": {"ms": round(cpu_s * 1e3, 1), "kind": "port", "cores": 1, "sample": "numpy restatement of Main_scaleLinear_rowcol, same frame"}}))
m.close()
