"""CPU tests of the oracle (the CPU restatement of the reference) against the hand-derived known-answer values of
SURVEY.md section 8(c), the committed golden fixtures, and internal consistency (numpy vs C restatement)."""
import math
import os

import numpy as np
import pytest

from oracle import anchors as OA, boxes as OB, localizer as OL, model as OM, nms as ON, nms_c
from oracle.rect import Rect

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _anchors(desc, cfg):
    return OA.Anchors(desc["layers"], desc["anchor_nets"], cfg["scales"])


def test_head_localizer_geometry_known_answers():
    # identical for vgg_small and vgg_large: head 1 stride 8 centre 12; heads 2/3/4 stride 16, 24/40/56
    for desc, rf in ((OM.VGG_SMALL, [50, 106, 138, 170]), (OM.VGG_LARGE, [60, 132, 164, 196])):
        a = _anchors(desc, OM.CFG_DUPLO)
        want = [(10, 8, 12), (13, 16, 24), (13, 16, 40), (13, 16, 56)] if desc is OM.VGG_SMALL else \
               [(12, 8, 12), (16, 16, 24), (16, 16, 40), (16, 16, 56)]
        for i, (nl, stride, centre) in enumerate(want):
            l = a.localizers[i]
            assert len(l.layers) == nl
            r0, r1 = l.featureToInputRect(0, 0, 1, 1), l.featureToInputRect(1, 1, 2, 2)
            assert r0.center() == (centre, centre)
            assert r1.center()[0] - r0.center()[0] == stride
            assert r0.width() == rf[i]


def test_anchor_lut_known_answers():
    a = _anchors(OM.VGG_SMALL, OM.CFG_DUPLO)
    want = {  # scale index -> three aspects' (min, max) at x = 1, float32
        0: [(-4, 28), (-10.627417, 34.62742), (0.6862915, 23.31371)],
        1: [(-8, 56), (-21.254833, 69.25484), (1.372583, 46.62742)],
        2: [(-24, 104), (-50.509666, 130.50967), (-5.254834, 85.25484)],
        3: [(-72, 184), (-125.01933, 237.01933), (-34.509666, 146.50967)],
    }
    for s, rows in want.items():
        for j, (mn, mx) in enumerate(rows):
            assert a.w[s, j, 0, 0] == np.float32(mn) and a.w[s, j, 0, 1] == np.float32(mx)
    assert a.w.dtype == np.float32 and a.w.shape == (4, 3, 200, 2)
    # the binary searches of findRangesXY assume monotone LUTs (Anchors.lua:87-104)
    assert np.all(np.diff(a.w[..., 0], axis=2) > 0) and np.all(np.diff(a.h[..., 1], axis=2) > 0)


def test_roi_localizer_known_answers():
    loc = OL.Localizer(OL.trunk_layer_info(OM.VGG_SMALL["layers"], 4))
    assert len(loc.layers) == 11
    cases = {(0, 0, 800, 450): (-2, -2, 52, 30), (100, 50, 300, 250): (4, 1, 21, 18),
             (100.5, 50.25, 300.75, 250.5): (4, 1, 21, 18), (0, 0, 16, 16): (-2, -2, 3, 3),
             (790, 440, 800, 450): (47, 25, 52, 30)}
    for src, dst in cases.items():
        assert loc.inputToFeatureRect(Rect(*src)).unpack() == tuple(float(v) for v in dst)
    big = OL.Localizer(OL.trunk_layer_info(OM.VGG_LARGE["layers"], 4))
    assert len(big.layers) == 14
    assert big.inputToFeatureRect(Rect(0, 0, 800, 450)).unpack() == (-3, -3, 53, 31)
    assert big.inputToFeatureRect(Rect(100, 50, 300, 250)).unpack() == (3, 0, 22, 19)


def test_anchor_counts_and_shapes():
    import torch
    p = OM.init_params(OM.VGG_SMALL, OM.CFG_DUPLO, seed=1)
    assert sum(v.numel() for v in p.values()) - 2 * 1024 == 26786154 - 2 * 1024  # 26.78 M floats
    with torch.no_grad():
        outs = OM.pnet_forward(OM.VGG_SMALL, p, torch.zeros(3, 122, 192))
    assert [tuple(o.shape) for o in outs] == [(18, 14, 22), (18, 6, 10), (18, 4, 8), (18, 2, 6), (384, 8, 12)]
    # full-size shapes by arithmetic: 450x800 -> blocks 225x400, 113x200, 57x100, 29x50; heads 55x98, 27x48, 25x46, 23x44
    h, w = 450, 800
    dims = []
    for _ in range(4):
        h, w = (h + 1) // 2, (w + 1) // 2
        dims.append((h, w))
    assert dims == [(225, 400), (113, 200), (57, 100), (29, 50)]
    heads = [(dims[2][0] - 2, dims[2][1] - 2), (dims[3][0] - 2, dims[3][1] - 2), (dims[3][0] - 4, dims[3][1] - 4),
             (dims[3][0] - 6, dims[3][1] - 6)]
    assert sum(a * b * 3 for a, b in heads) == 26544


def test_rect_semantics():
    a, b = Rect(0, 0, 10, 10), Rect(5, 5, 15, 15)
    assert Rect.IoU(a, b) == 25 / 175  # no +1 (Rect.lua:138-141)
    assert a.overlaps(b) and not a.overlaps(Rect(10, 0, 20, 10))  # strict
    assert Rect.intersect(a, Rect(20, 20, 30, 30)).unpack() == (0, 0, 0, 0)
    assert Rect(-1.5, 0.2, 3.1, 4).snapToInt().unpack() == (-2, 0, 4, 4)
    assert Rect(-5, -5, 5, 50).clip(Rect(0, 0, 10, 10)).unpack() == (0, 0, 5, 10)
    t = OA.Anchors.inputToAnchor(Rect(0, 0, 10, 20), Rect(5, 10, 10, 30))
    r = OA.Anchors.anchorToInput(Rect(0, 0, 10, 20), t)
    assert np.allclose(r.unpack(), (5, 10, 10, 30), atol=1e-5)


# ------------------------------------------------------------------------------------------------ NMS
def test_nms_hand_cases():
    # two identical boxes: the higher index is picked first (tie rule), the other is suppressed
    b = np.array([[0, 0, 9, 9], [0, 0, 9, 9]], np.float32)
    assert ON.nms(b, 0.5).tolist() == [1]
    # disjoint boxes: all kept, ordered by y2 descending
    b = np.array([[0, 0, 9, 9], [20, 20, 29, 29], [40, 5, 49, 14]], np.float32)
    assert ON.nms(b, 0.1).tolist() == [1, 2, 0]
    # +1 pixel convention (nms.lua:35): touching boxes overlap by one pixel column
    b = np.array([[0, 0, 9, 9], [9, 0, 18, 10]], np.float32)
    inter, union = 1 * 10, 100 + 110 - 10
    assert ON.nms(b, inter / union + 1e-6).tolist() == [1, 0]
    assert ON.nms(b, inter / union - 1e-6).tolist() == [1]
    # `le`: IoU == overlap is kept (nms.lua:96)
    b = np.array([[0, 0, 3, 3], [0, 0, 3, 7]], np.float32)  # inter 16, union 32 -> 0.5 exactly
    assert ON.nms(b, 0.5).tolist() == [1, 0]
    # Q1: a score tensor is ignored -> order by y2; a number selects the column; 'area' the area
    b = np.array([[0, 0, 9, 9, 0.9], [20, 20, 29, 29, 0.1]], np.float32)
    assert ON.nms(b, 0.5, b[:, 4]).tolist() == [1, 0]
    assert ON.nms(b, 0.5, 5).tolist() == [0, 1]
    assert ON.nms(np.zeros((0, 4), np.float32), 0.5).tolist() == []


@pytest.mark.parametrize("n", [1, 2, 3, 33, 257, 1000, 5000])
@pytest.mark.parametrize("thr", [0.1, 0.25])
def test_nms_numpy_vs_c(n, thr):
    b = OB.sweep_boxes(n, seed=n)
    for mode, arg in ((0, None), (1, "area"), (2, 1)):
        assert np.array_equal(ON.nms(b, thr, arg), nms_c.nms(b, thr, mode, 0))


def test_nms_ties_and_properties():
    rng = np.random.default_rng(3)
    b = OB.sweep_boxes(3000, seed=3)
    b[:, 3] = np.round(b[:, 3] / 8) * 8  # many ties in the y2 key
    b[:, 1] = np.minimum(b[:, 1], b[:, 3] - 1)
    p = ON.nms(b, 0.25)
    assert np.array_equal(p, nms_c.nms(b, 0.25))
    keys = b[p, 3]
    assert np.all(np.diff(keys) <= 0)  # pick order = key descending
    assert np.array_equal(np.sort(ON.nms(b[p], 0.25)), np.arange(len(p)))  # idempotent: no pick suppresses another
    tf = OB.sweep_boxes(2000, seed=5, tie_free=True)
    assert len(np.unique(tf[:, 3])) == 2000


def test_nms_segmented_matches_loop():
    b = OB.sweep_boxes(4000, seed=7)
    perm, seg = OB.class_segments(4000, 21, seed=7)
    b = b[perm]
    p, c = nms_c.nms_segmented(b, seg, 0.1, threads=4)
    p2, c2 = ON.nms_segmented(b, seg, 0.1)
    assert np.array_equal(c, c2)
    off = np.concatenate([[0], np.cumsum(c2)])
    for s in range(21):
        assert np.array_equal(p[seg[s]:seg[s] + c[s]], p2[off[s]:off[s + 1]])


def test_golden_fixtures():
    """The committed fixtures were produced by tests/golden/make_golden.py from this oracle; they pin the oracle
    against silent drift (the reference itself ships no vectors: parity unpinned, see oracle/__init__.py)."""
    g = np.load(os.path.join(GOLD, "geometry_nms.npz"))
    a = _anchors(OM.VGG_SMALL, OM.CFG_DUPLO)
    assert np.array_equal(a.w, g["small_w"]) and np.array_equal(a.h, g["small_h"])
    al = _anchors(OM.VGG_LARGE, OM.CFG_IMAGENET)
    assert np.array_equal(al.w, g["large_w"]) and np.array_equal(al.h, g["large_h"])
    loc = OL.Localizer(OL.trunk_layer_info(OM.VGG_SMALL["layers"], 4))
    got = np.array([loc.inputToFeatureRect(Rect(*r)).unpack() for r in g["roi_in"]])
    assert np.array_equal(got, g["roi_out_small"])
    b = g["nms_boxes"]
    assert np.array_equal(ON.nms(b, 0.25), g["nms_pick_025"])
    assert np.array_equal(ON.nms(b, 0.1), g["nms_pick_010"])
    assert np.array_equal(ON.nms(b, 0.25, "area"), g["nms_pick_area"])


def test_host_mirror_anchor_lookups_match_oracle():
    """Anchors:findRangesXY / Anchors:findNearby of the host mirror (host-side logic, like the reference's) against the
    oracle on a host-only context (no GPU needed): same ranges, same anchors, same order."""
    import frcnn_b200 as F
    m = F.vgg_small(F.duplo_cfg, device=-1)
    ga = F.Anchors(m)
    oa = OA.Anchors(OM.VGG_SMALL["layers"], OM.VGG_SMALL["anchor_nets"], OM.CFG_DUPLO["scales"])
    assert np.array_equal(ga.w, oa.w) and np.array_equal(ga.h, oa.h)
    rng = np.random.default_rng(4)
    img = Rect(0, 0, 800, 450)
    for _ in range(20):
        bw, bh = rng.uniform(5, 400), rng.uniform(5, 300)
        x, y = rng.uniform(-40, 800 - bw), rng.uniform(-40, 450 - bh)
        r = Rect(x, y, x + bw, y + bh)
        for clip in (None, img):
            a, b = ga.findRangesXY(r, clip), oa.findRangesXY(r, clip)
            assert [(q["layer"], q["aspect"], q["lx"], q["ly"], q["ux"], q["uy"]) for q in a] == \
                   [(q["layer"], q["aspect"], q["lx"], q["ly"], q["ux"], q["uy"]) for q in b]
        cx, cy = rng.uniform(0, 800), rng.uniform(0, 450)
        fa, fb = ga.findNearby(cx, cy), oa.findNearby(cx, cy)
        assert [(q.layer, q.aspect, q.index[1], q.index[2]) for q in fa] == [(q.layer, q.aspect, q.index[1], q.index[2]) for q in fb]
    assert len(ga.findNearby(400.0, 225.0)) > 0
    m.close()
