"""Precision contract of the bf16 tensor-core path against the PURE fp32 restatement of the reference, measured end to
end from the frame.  Test infrastructure only: imported by tests/ and by bench.py's
cpu_baseline leg, never by the product path.

The reference computes pnet / cnet in fp32 (cunn SGEMM); the CUDA path rounds conv / Linear operands to bf16 and
accumulates in fp32.  This module states what that costs where the reference takes DISCRETE decisions:
  * per output map: max / mean absolute error in the map's own units (logits for the 18-channel heads);
  * the decision variable of Detector.lua:52-54, d = c_fg - c_bg (p > 0.95 <=> d > ln 19): its error, and how many
    anchors sit closer to the threshold than the observed error (the anchors that CAN flip);
  * Jaccard index of the match list (Detector.lua:54-59), of the candidates after nms(bb, 0.25) (Detector.lua:82) and
    of the winners by (class, anchor) (Detector.lua:107-136);
  * for the winners both paths agree on: largest difference of the refined box r2 relative to the box size, and of
    the class log-probability."""
import math

import numpy as np

from . import detector as OD
from . import model as OM

LN19 = math.log(0.95 / 0.05)


def anchor_key(x):
    a = x["a"]
    return (x["l"], a.aspect, a.index[1], a.index[2])


def winner_key(x):
    return (x["class"],) + anchor_key(x)


def jaccard(a, b):
    a, b = set(a), set(b)
    u = len(a | b)
    return 1.0 if u == 0 else len(a & b) / u


def map_errors(got, want):
    """got / want: lists of 5 arrays (4 head maps [18][h][w], feature map [C][h][w])."""
    out = []
    for i, (g, w) in enumerate(zip(got, want)):
        g = np.asarray(g, dtype=np.float64)
        w = np.asarray(w, dtype=np.float64)
        e = np.abs(g - w)
        rec = dict(map="head%d" % (i + 1) if i < 4 else "feature", max_abs=float(e.max()), mean_abs=float(e.mean()),
                   ref_rms=float(np.sqrt((w * w).mean())), ref_max=float(np.abs(w).max()))
        if i < 4:
            # decision variable of Detector.lua:52-54 per anchor: d = fg - bg logit, threshold ln(19)
            dg = np.stack([g[6 * a] - g[6 * a + 1] for a in range(3)])
            dw = np.stack([w[6 * a] - w[6 * a + 1] for a in range(3)])
            de = np.abs(dg - dw)
            rec.update(margin_max_abs=float(de.max()), margin_mean_abs=float(de.mean()),
                       anchors=int(dw.size), anchors_within_err_of_threshold=int((np.abs(dw - LN19) <= de.max()).sum()),
                       flips=int(((dg > LN19) != (dw > LN19)).sum()))
            # box regression channels (x, y, w, h): what Anchors.anchorToInput consumes
            reg = np.stack([np.abs(g[6 * a + 2:6 * a + 6] - w[6 * a + 2:6 * a + 6]) for a in range(3)])
            rec.update(reg_max_abs=float(reg.max()), reg_mean_abs=float(reg.mean()))
        out.append(rec)
    return out


def stage_report(got_inter, got_winners, want_inter, want_winners):
    """got_*: the discrete stages computed from the CUDA path's maps; want_*: the pure fp32 oracle's."""
    gm, wm = [anchor_key(x) for x in got_inter["matches"]], [anchor_key(x) for x in want_inter["matches"]]
    gc, wc = [anchor_key(x) for x in got_inter["candidates"]], [anchor_key(x) for x in want_inter["candidates"]]
    gw = {winner_key(x): x for x in got_winners}
    ww = {winner_key(x): x for x in want_winners}
    common = sorted(set(gw) & set(ww))
    box_rel, conf_abs = 0.0, 0.0
    for k in common:
        a, b = gw[k], ww[k]
        size = max(b["r2"].width(), b["r2"].height(), 1.0)
        box_rel = max(box_rel, float(np.max(np.abs(np.array(a["r2"].unpack()) - np.array(b["r2"].unpack()))) / size))
        conf_abs = max(conf_abs, abs(float(a["confidence"]) - float(b["confidence"])))
    # winners that agree on the anchor but not on the class
    # geometric agreement (independent of WHICH anchor of a cluster survived the bottom-edge-ordered greedy NMS): share of
    # the fp32 winners that have a CUDA winner of the same class overlapping them with IoU >= 0.7, and vice versa
    def iou(a, b):  # Rect.lua:126-141 on the unpacked corners (the two sides carry different Rect classes)
        ax0, ay0, ax1, ay1 = a.unpack()
        bx0, by0, bx1, by1 = b.unpack()
        w, h = min(ax1, bx1) - max(ax0, bx0), min(ay1, by1) - max(ay0, by0)
        i = w * h if w > 0 and h > 0 else 0.0
        u = (ax1 - ax0) * (ay1 - ay0) + (bx1 - bx0) * (by1 - by0) - i
        return i / u if u > 0 else 0.0

    def covered(src, dst):
        n = 0
        for k, a in src.items():
            for k2, b in dst.items():
                if k2[0] == k[0] and iou(a["r2"], b["r2"]) >= 0.7:
                    n += 1
                    break
        return n
    cov_w = covered(ww, gw) / max(len(ww), 1) if ww else 1.0
    cov_g = covered(gw, ww) / max(len(gw), 1) if gw else 1.0
    ga = {k[1:]: k[0] for k in gw}
    wa = {k[1:]: k[0] for k in ww}
    class_flips = sum(1 for k in set(ga) & set(wa) if ga[k] != wa[k])
    return dict(matches=dict(cuda=len(gm), fp32=len(wm), jaccard=jaccard(gm, wm)),
                candidates=dict(cuda=len(gc), fp32=len(wc), jaccard=jaccard(gc, wc)),
                winners=dict(cuda=len(gw), fp32=len(ww), common=len(common), jaccard=jaccard(gw, ww),
                             class_flips_on_common_anchor=class_flips, r2_max_rel_to_box=box_rel,
                             confidence_max_abs=conf_abs, fp32_covered_by_cuda=cov_w, cuda_covered_by_fp32=cov_g))


def precision_report(desc, cfg, params, img, cuda_maps, cuda_winners, fp32_result=None, quant=None):
    """cuda_maps: the 5 outputs of the CUDA pnet:forward (CPU tensors); cuda_winners: what the CUDA Detector:detect
    returned for the same frame.  The stages of the CUDA side between the maps and the winners (matches, candidates)
    are taken from the oracle's discrete code run on the CUDA maps -- the GPU tests assert those are bit-identical
    to what the kernels produce.  fp32_result: (winners, intermediates) of a previous pure-fp32 oracle run on `img`."""
    if fp32_result is None:
        fp32_result = OD.Detector(desc, cfg, params).detect(img, return_intermediates=True)
    want, want_inter = fp32_result
    _, got_inter = OD.Detector(desc, cfg, params, quant=quant or OM.fp16_round, quant_heads=None).detect(
        img, outputs=cuda_maps, return_intermediates=True)
    want_winners = [x for c in sorted(want) for x in want[c]]
    rep = stage_report(got_inter, cuda_winners, want_inter, want_winners)
    rep["maps"] = map_errors([np.asarray(t) for t in cuda_maps], [t.numpy() for t in want_inter["outputs"]])
    return rep
