"""Restatement of Detector.lua:17-141 and objective.lua:5-13 (ROI crop).  Test infrastructure only."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .anchors import Anchors
from .localizer import Localizer, trunk_layer_info
from .nms import nms
from .rect import Rect
from . import model as M


def logsoftmax2_fg(c1, c2):
    """nn.LogSoftMax on the (fg, bg) pair, element 1 (Detector.lua:52).  The cunn fp32 kernel's exact
    rounding is unpinned; the oracle evaluates the 2-element log-softmax in double and rounds once to fp32
    (<= 1 ulp from any faithful fp32 implementation)."""
    c1 = np.asarray(c1, dtype=np.float64)
    c2 = np.asarray(c2, dtype=np.float64)
    m = np.maximum(c1, c2)
    return ((c1 - m) - np.log(np.exp(c1 - m) + np.exp(c2 - m))).astype(np.float32)


def roi_crop_index(rect, localizer, fh, fw):
    """extract_roi_pooling_input (objective.lua:5-13): 0-based half-open (y0, y1, x0, x1) crop of the feature
    map, or None where the reference would raise (clipped max == 0)."""
    r = localizer.inputToFeatureRect(rect)
    r = r.clip(Rect(0, 0, fw, fh))
    y_lo, y_hi = min(r.minY + 1, r.maxY), r.maxY  # 1-based inclusive
    x_lo, x_hi = min(r.minX + 1, r.maxX), r.maxX
    if y_lo < 1 or x_lo < 1 or y_hi < y_lo or x_hi < x_lo:
        return None
    return int(y_lo) - 1, int(y_hi), int(x_lo) - 1, int(x_hi)


def roi_pool(fmap, rect, localizer, kh=6, kw=6):
    """amp:forward(crop):view(kh*kw*C) (Detector.lua:96-97): adaptive max pool, channel-major flatten."""
    C, fh, fw = fmap.shape
    idx = roi_crop_index(rect, localizer, fh, fw)
    if idx is None:
        raise IndexError("empty ROI crop (the reference raises here, SURVEY Q8)")
    y0, y1, x0, x1 = idx
    out, ind = F.adaptive_max_pool2d(fmap[:, y0:y1, x0:x1].unsqueeze(0), (kh, kw), return_indices=True)
    return out.reshape(-1), ind.reshape(-1), idx


def decode(outputs, anchors, input_rect, threshold=0.95):
    """Detector.lua:36-66: ordered (layer, y, x, aspect) list of matches {p, a, r, l}."""
    matches = []
    for i in range(4):
        layer = outputs[i].numpy()
        _, H, W = layer.shape
        lp = np.stack([logsoftmax2_fg(layer[6 * a], layer[6 * a + 1]) for a in range(3)], axis=-1)  # [H][W][3]
        keep = np.exp(lp.astype(np.float64)) > threshold  # math.exp(c[1]) > 0.95 in double
        ys, xs, as_ = np.nonzero(keep)  # C order == (y, x, a) loop order of Detector.lua:42-45
        for y, x, a in zip(ys, xs, as_):
            anc = anchors.get(i + 1, a + 1, y + 1, x + 1)
            r = Anchors.anchorToInput(anc, layer[6 * a + 2:6 * a + 6, y, x])
            if r.overlaps(input_rect):
                matches.append(dict(p=lp[y, x, a], a=anc, r=r, l=i + 1))
    return matches


class Detector:
    def __init__(self, desc, cfg, params, dropout_eval_scale=None, quant=None, quant_heads="same"):  # Detector.lua:8-15
        self.desc, self.cfg, self.p = desc, cfg, params
        self.anchors = Anchors(desc["layers"], desc["anchor_nets"], cfg["scales"])
        self.localizer = Localizer(trunk_layer_info(desc["layers"], len(desc["layers"])))
        self.dropout_eval_scale = dropout_eval_scale
        self.quant = quant
        self.quant_heads = quant_heads

    def detect(self, img, outputs=None, return_intermediates=False):  # Detector.lua:17-141
        cfg = self.cfg
        kh, kw = cfg["roi_pooling"]["kh"], cfg["roi_pooling"]["kw"]
        bgclass = cfg["class_count"] + 1
        input_rect = Rect(0, 0, img.shape[2], img.shape[1])
        with torch.no_grad():
            if outputs is None:
                outputs = M.pnet_forward(self.desc, self.p, img, dropout_eval_scale=self.dropout_eval_scale,
                                         quant=self.quant)
            matches = decode(outputs, self.anchors, input_rect)
            inter = dict(outputs=outputs, matches=matches, candidates=[], cinput=None, coutputs=None)
            winners = {}
            if len(matches) > 0:
                bb = np.stack([m["r"].totensor() for m in matches])  # Detector.lua:74-79
                pick = nms(bb, 0.25, None)  # score tensor is ignored (Q1) -> key = y2
                candidates = [matches[i] for i in pick]
                fmap = outputs[4]
                cinput = torch.stack([roi_pool(fmap, v["r"], self.localizer, kh, kw)[0] for v in candidates])
                bbox_out, cls_out = M.cnet_forward(self.desc, self.p, cinput, quant=self.quant, quant_heads=self.quant_heads)
                inter.update(candidates=candidates, cinput=cinput, coutputs=(bbox_out, cls_out), pick=pick)
                yclass = {}
                for i, x in enumerate(candidates):  # Detector.lua:106-122
                    x["r2"] = Anchors.anchorToInput(x["r"], bbox_out[i].numpy())
                    cprob = cls_out[i].numpy()
                    c = int(np.argmax(cprob))  # torch.sort(cprob, 1, true)[1]; first max on ties
                    x["class"] = c + 1
                    x["confidence"] = cprob[c]
                    if x["class"] != bgclass and math.exp(float(x["confidence"])) > 0.2:
                        yclass.setdefault(x["class"], []).append(x)
                for c, lst in yclass.items():  # Detector.lua:125-136 (pairs() order unspecified, Q7)
                    bb = np.zeros((len(lst), 5), dtype=np.float32)
                    for j, r in enumerate(lst):
                        bb[j, :4] = r["r2"].totensor()
                        bb[j, 4] = r["confidence"]
                    pk = nms(bb, 0.1, None)  # bb[{{},5}] is a tensor -> ignored (Q1)
                    winners[c] = [lst[i] for i in pk]
        if return_intermediates:
            return winners, inter
        return winners
