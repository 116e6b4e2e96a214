"""GPU parity tests of the NMS kernels (csrc/nms.cu) through the C ABI against the oracle (oracle/nms.py numpy
restatement of nms.lua, oracle/nms_ref.c C restatement).  Bar: bit-exact pick indices in pick order."""
import os

import numpy as np
import pytest

from oracle import boxes as OB, nms as ON, nms_c

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_hand_cases(F):
    b = np.array([[0, 0, 9, 9], [0, 0, 9, 9]], np.float32)
    assert F.nms(b, 0.5).tolist() == [1]
    b = np.array([[0, 0, 9, 9], [20, 20, 29, 29], [40, 5, 49, 14]], np.float32)
    assert F.nms(b, 0.1).tolist() == [1, 2, 0]
    b = np.array([[0, 0, 9, 9], [9, 0, 18, 10]], np.float32)
    inter, union = 1 * 10, 100 + 110 - 10
    assert F.nms(b, inter / union + 1e-6).tolist() == [1, 0]
    assert F.nms(b, inter / union - 1e-6).tolist() == [1]
    b = np.array([[0, 0, 3, 3], [0, 0, 3, 7]], np.float32)  # IoU == 0.5 exactly: `le` keeps it (nms.lua:96)
    assert F.nms(b, 0.5).tolist() == [1, 0]
    b = np.array([[0, 0, 9, 9, 0.9], [20, 20, 29, 29, 0.1]], np.float32)
    assert F.nms(b, 0.5, b[:, 4]).tolist() == [1, 0]  # Q1: score tensor ignored
    assert F.nms(b, 0.5, 5).tolist() == [0, 1]
    assert F.nms(np.zeros((0, 4), np.float32), 0.5).tolist() == []
    assert F.nms(np.array([[1, 2, 3, 4]], np.float32), 0.5).tolist() == [0]


def test_golden(F):
    g = np.load(os.path.join(GOLD, "geometry_nms.npz"))
    b = g["nms_boxes"]
    assert np.array_equal(F.nms(b, 0.25), g["nms_pick_025"])
    assert np.array_equal(F.nms(b, 0.1), g["nms_pick_010"])
    assert np.array_equal(F.nms(b, 0.25, "area"), g["nms_pick_area"])


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 257, 1000, 1023, 1024, 1025, 2049, 5000, 8192])
@pytest.mark.parametrize("thr", [0.1, 0.25, 0.7])
def test_vs_oracle_small_path(F, n, thr):
    b = OB.sweep_boxes(n, seed=n)
    for mode, arg in ((0, None), (1, "area"), (2, 1)):
        assert np.array_equal(F.nms(b, thr, arg), nms_c.nms(b, thr, mode, 0)), (n, thr, mode)


@pytest.mark.parametrize("n", [8193, 20000, 100000])
def test_vs_oracle_radix_path(F, n):
    b = OB.sweep_boxes(n, seed=n)
    for thr in (0.1, 0.25):
        assert np.array_equal(F.nms(b, thr), nms_c.nms(b, thr))
    assert np.array_equal(F.nms(b, 0.25, "area"), nms_c.nms(b, 0.25, 1, 0))


def test_ties(F):
    for n in (3000, 30000):
        b = OB.sweep_boxes(n, seed=3)
        b[:, 3] = np.round(b[:, 3] / 8) * 8  # massive ties in the y2 key: order defined as (key, index)
        b[:, 1] = np.minimum(b[:, 1], b[:, 3] - 1)
        p = F.nms(b, 0.25)
        assert np.array_equal(p, nms_c.nms(b, 0.25))
        assert np.all(np.diff(b[p, 3]) <= 0)
    b = np.tile(np.array([[5, 5, 50, 50]], np.float32), (2000, 1))  # all identical
    assert F.nms(b, 0.5).tolist() == [1999]
    # disjoint boxes: nothing suppressed, several rounds, pick order = key descending then index descending
    gx, gy = np.meshgrid(np.arange(60), np.arange(50))
    b = np.stack([gx.ravel() * 20, gy.ravel() * 20, gx.ravel() * 20 + 10, gy.ravel() * 20 + 10], 1).astype(np.float32)
    p = F.nms(b, 0.1)
    assert len(p) == 3000 and np.array_equal(p, nms_c.nms(b, 0.1))


def test_special_values(F):
    b = OB.sweep_boxes(500, seed=9)
    b[7] = [10, 10, 5, 5]          # inverted box: area (-4)*(-4) = 16, intersections clamp to 0
    b[11] = [0, 0, -1, -1]         # zero area -> 0/0 = NaN against itself-like boxes
    b[13, 2] = np.inf
    b[17] = np.nan                 # NaN IoU: `iou <= thr` false -> dropped without being picked unless reached first
    got, want = F.nms(b, 0.25, "area"), nms_c.nms(b, 0.25, 1, 0)
    # NaN keys have no defined sort position in either implementation: compare on the NaN-free prefix case only
    b2 = b.copy()
    b2[17] = [3, 3, 8, 8]
    assert np.array_equal(F.nms(b2, 0.25), nms_c.nms(b2, 0.25))
    assert np.array_equal(F.nms(b2, 0.25, "area"), nms_c.nms(b2, 0.25, 1, 0))
    assert len(got) > 0 and len(want) > 0


def test_zero_intersection_paths(F):
    """Disjoint pairs never reach the divider in the kernels: the comparison result of +0 / d is derived instead.
    Zero-area pairs (0 / 0 = NaN -> dropped), negative areas, and thresholds <= 0 must match the dividing oracle."""
    rng = np.random.default_rng(21)
    b = OB.sweep_boxes(1500, seed=21)
    z = rng.choice(1500, 200, replace=False)
    b[z[:100], 2] = b[z[:100], 0] - 1       # width 0: area 0, every intersection 0
    b[z[100:150], 3] = b[z[100:150], 1] - 3  # negative height: negative area
    b[z[150:], :] = b[z[150], :]             # exact duplicates of one box
    for thr in (0.25, 0.0, -0.5, 1.5):
        for arg, mode in ((None, 0), ("area", 1)):
            assert np.array_equal(F.nms(b, thr, arg), nms_c.nms(b, thr, mode, 0)), (thr, mode)
    zero = np.zeros((300, 4), np.float32)
    zero[:, 2:] = -1                          # all areas 0, all sums of areas 0: every test is 0 / 0
    zero[:, 3] -= np.arange(300)
    zero[:, 1] = zero[:, 3] + 1
    assert np.array_equal(F.nms(zero, 0.25), nms_c.nms(zero, 0.25))


@pytest.mark.parametrize("n,n_seg", [(4000, 21), (50, 21), (64000, 21), (300000, 21)])
def test_segmented(F, n, n_seg):
    b = OB.sweep_boxes(n, seed=7)
    perm, seg = OB.class_segments(n, n_seg, seed=7)
    b = b[perm]
    for thr in (0.1, 0.25):
        p, c = F.nms_segmented(b, seg, thr)
        p2, c2 = nms_c.nms_segmented(b, seg, thr, threads=8)
        assert np.array_equal(c, c2)
        for s in range(n_seg):
            assert np.array_equal(p[seg[s]:seg[s] + c[s]], p2[seg[s]:seg[s] + c2[s]])


def test_segmented_empty_segments(F):
    b = OB.sweep_boxes(100, seed=1)
    seg = np.array([0, 0, 40, 40, 100, 100], np.int64)
    p, c = F.nms_segmented(b, seg, 0.25)
    p2, c2 = nms_c.nms_segmented(b, seg, 0.25)
    assert np.array_equal(c, c2) and c[0] == 0 and c[2] == 0 and c[4] == 0
    for s in range(5):
        assert np.array_equal(p[seg[s]:seg[s] + c[s]], p2[seg[s]:seg[s] + c2[s]])


def test_row_stride_and_device_entry(F):
    import torch
    b = np.concatenate([OB.sweep_boxes(3000, seed=2), np.random.default_rng(0).random((3000, 3), np.float32)], 1)
    assert np.array_equal(F.nms(b, 0.25, 6), nms_c.nms(b, 0.25, 2, 5))
    bd = torch.from_numpy(b).cuda()
    p, c = F.nms_segmented_dev(bd, np.array([0, 3000]), 0.25)
    assert np.array_equal(p[:int(c[0])].cpu().numpy(), nms_c.nms(b, 0.25))


def test_full_size_properties(F):
    """BASELINE config 5 at 1M boxes: size-independent properties (sortedness of the pick keys, idempotence, and
    agreement of a checksum of the picks with the C oracle on the 21-class split)."""
    n = 1_000_000
    b = OB.sweep_boxes(n, seed=0)
    perm, seg = OB.class_segments(n, 21, seed=0)
    b = b[perm]
    p, c = F.nms_segmented(b, seg, 0.25)
    p2, c2 = nms_c.nms_segmented(b, seg, 0.25, threads=8)
    assert np.array_equal(c, c2)
    for s in range(21):
        mine = p[seg[s]:seg[s] + c[s]]
        assert np.array_equal(mine, p2[seg[s]:seg[s] + c2[s]])
        keys = b[seg[s] + mine, 3]
        assert np.all(np.diff(keys) <= 0)
        kept = b[seg[s] + mine]
        assert np.array_equal(np.sort(F.nms(kept, 0.25)), np.arange(len(kept)))  # no pick suppresses another
