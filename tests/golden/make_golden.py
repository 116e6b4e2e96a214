"""Generates the golden fixtures from the oracle (run from the repo root: python tests/golden/make_golden.py).
The reference cannot be executed here (no Lua / Torch7), so these vectors come from the CPU restatement, which is
itself pinned to the hand-derived known answers of SURVEY.md 8(c) by tests/test_oracle.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import anchors as OA, boxes as OB, localizer as OL, model as OM, nms as ON  # noqa: E402
from oracle.detector import decode  # noqa: E402
from oracle.rect import Rect  # noqa: E402

out = os.path.dirname(os.path.abspath(__file__))
small = OA.Anchors(OM.VGG_SMALL["layers"], OM.VGG_SMALL["anchor_nets"], OM.CFG_DUPLO["scales"])
large = OA.Anchors(OM.VGG_LARGE["layers"], OM.VGG_LARGE["anchor_nets"], OM.CFG_IMAGENET["scales"])
rng = np.random.default_rng(0)
roi_in = np.concatenate([
    np.array([[0, 0, 800, 450], [100, 50, 300, 250], [100.5, 50.25, 300.75, 250.5], [0, 0, 16, 16],
              [790, 440, 800, 450], [-30.5, -12.25, 40.75, 33.5], [700.2, 400.9, 905.5, 512.1]], dtype=np.float64),
    np.stack([rng.uniform(-50, 780, 64), rng.uniform(-50, 430, 64), np.zeros(64), np.zeros(64)], axis=1)])
roi_in[7:, 2] = roi_in[7:, 0] + rng.uniform(1, 400, 64)
roi_in[7:, 3] = roi_in[7:, 1] + rng.uniform(1, 300, 64)
loc_s = OL.Localizer(OL.trunk_layer_info(OM.VGG_SMALL["layers"], 4))
loc_l = OL.Localizer(OL.trunk_layer_info(OM.VGG_LARGE["layers"], 4))
boxes = OB.sweep_boxes(2000, seed=11)
np.savez_compressed(
    os.path.join(out, "geometry_nms.npz"),
    small_w=small.w, small_h=small.h, large_w=large.w, large_h=large.h, roi_in=roi_in,
    roi_out_small=np.array([loc_s.inputToFeatureRect(Rect(*r)).unpack() for r in roi_in]),
    roi_out_large=np.array([loc_l.inputToFeatureRect(Rect(*r)).unpack() for r in roi_in]),
    nms_boxes=boxes, nms_pick_025=ON.nms(boxes, 0.25), nms_pick_010=ON.nms(boxes, 0.1),
    nms_pick_area=ON.nms(boxes, 0.25, "area"))

# RPN decode fixture: small seeded head maps (90x160 input -> heads 10x18, 4x8, 2x6; layer 4 is empty at this size,
# so a 122x192 input is used: blocks 61x96, 31x48, 16x24, 8x12 -> heads 14x22, 6x10, 4x8, 2x6)
import torch  # noqa: E402
g = torch.Generator().manual_seed(5)
dims = [(14, 22), (6, 10), (4, 8), (2, 6)]
heads = [torch.randn((18,) + d, generator=g) * 1.5 for d in dims]
for h in heads:
    h[0::6] += 1.0
m = decode(heads, small, Rect(0, 0, 192, 122))
np.savez_compressed(
    os.path.join(out, "decode.npz"), **{"head%d" % i: h.numpy() for i, h in enumerate(heads)},
    anchor=np.array([[x["l"], x["a"].aspect, x["a"].index[1], x["a"].index[2]] for x in m], dtype=np.int32),
    logp=np.array([x["p"] for x in m], dtype=np.float32), r=np.array([x["r"].unpack() for x in m], dtype=np.float64),
    box=np.stack([x["r"].totensor() for x in m]))
print("golden fixtures written:", len(m), "decode matches")

# ---- anchor labelling, optimiser step, frame normalisation (next_rows.npz)
from oracle import optim as OO, preprocess as OP  # noqa: E402
rng = np.random.default_rng(3)
rois = []
for _ in range(6):
    bw, bh = rng.uniform(20, 300), rng.uniform(20, 250)
    x, y = rng.uniform(0, 800 - bw), rng.uniform(0, 450 - bh)
    rois.append([x, y, x + bw, y + bh])
rois[0] = [100, 100, 112, 109]          # too small for any positive anchor: the best-set branch
rois = np.array(rois, dtype=np.float64)
roi_list = [{"rect": Rect(*r)} for r in rois]
img = Rect(0, 0, 800, 450)
pos = small.findPositive(roi_list, img, 0.6, 0.3, True)
idx = {id(r): i for i, r in enumerate(roi_list)}
rnd = rng.integers(0, 2 ** 32, 3 * 600, dtype=np.uint64).astype(np.uint32)
neg, trials = small.sampleNegative(img, roi_list, 0.3, 96, rnd)
w = rng.standard_normal(4099).astype(np.float32)
gsteps = [(rng.standard_normal(4099) * 10.0 ** rng.integers(-5, 2, 4099)).astype(np.float32) for _ in range(3)]
w_ref, st = w.copy(), {}
for gstep in gsteps:
    OO.rmsprop_step(w_ref, OO.gradient_div(gstep, 37.0), st, learningRate=1e-3, alpha=0.99, epsilon=1e-8, weightDecay=0.0005)
frame = rng.uniform(0, 1, (3, 45, 80)).astype(np.float32)
frame[:, 10:25, 20:50] += 0.7
np.savez_compressed(
    os.path.join(out, "next_rows.npz"), rois=rois,
    pos=np.array([[a.layer, a.aspect, a.index[1], a.index[2], idx[id(r)]] for a, r in pos], dtype=np.int32),
    rnd=rnd, neg=np.array([[a.layer, a.aspect, a.index[1], a.index[2]] for (a,) in neg], dtype=np.int32), neg_trials=np.int32(trials),
    opt_w0=w, opt_g=np.stack(gsteps), opt_w3=w_ref, opt_m3=st["m"],
    frame=frame, frame_norm=OP.normalize_frame(torch.from_numpy(frame), rgb_to_yuv=True).numpy())
print("next_rows fixture:", len(pos), "positives,", len(neg), "negatives in", trials, "trials")
