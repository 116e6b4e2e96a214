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
