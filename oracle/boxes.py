"""Synthetic NMS workloads of BASELINE.md section 3 config 5.  Test / bench infrastructure only."""
import numpy as np


def sweep_boxes(n, seed=0, tie_free=False, width=800.0, height=450.0):
    """fp32 boxes {x1,y1,x2,y2}: centres U([0,W]x[0,H]), side exp(U(ln32, ln256)), aspect {1,2,1/2}*exp(N(0,.1))."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(0, width, n)
    cy = rng.uniform(0, height, n)
    s = np.exp(rng.uniform(np.log(32.0), np.log(256.0), n))
    asp = rng.choice([1.0, 2.0, 0.5], n) * np.exp(rng.normal(0, 0.1, n))
    w = s * np.sqrt(asp)
    h = s / np.sqrt(asp)
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1).astype(np.float32)
    if tie_free:
        # make every y2 distinct in fp32 while keeping the distribution: sort, then nudge duplicates upward
        order = np.argsort(b[:, 3], kind="stable")
        y2 = b[order, 3].copy()
        for i in range(1, n):
            if y2[i] <= y2[i - 1]:
                y2[i] = np.nextafter(y2[i - 1], np.float32(np.inf), dtype=np.float32)
        b[order, 3] = y2
    return b


def class_segments(n, n_seg=21, seed=0):
    """Class id uniform in 0..n_seg-1; returns (permutation grouping boxes by class, seg_offsets)."""
    rng = np.random.default_rng(seed + 1)
    cls = rng.integers(0, n_seg, n)
    perm = np.argsort(cls, kind="stable")
    counts = np.bincount(cls, minlength=n_seg)
    return perm, np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
