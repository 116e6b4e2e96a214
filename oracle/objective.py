"""Restatement of objective.lua:45-218 (lossAndGradient) for one image on PyTorch-CPU with autograd.  Test
infrastructure only.

The reference applies every criterion's backward directly (objective.lua:104-114,132-134,171-179), which is the
gradient of   sum CE(2-vec) + 10 * sum SmoothL1(reg)   +   10 * SmoothL1_sum(crout_pos) + mean NLL(ccout)
with the regressed proposal (`reg_proposal`, objective.lua:111) treated as a constant and the bbox outputs of the
negative examples zeroed before the loss (objective.lua:170).  Un-vendored criteria restated from torch7 `nn`:
CrossEntropyCriterion = LogSoftMax + ClassNLL; SmoothL1Criterion(sizeAverage=false) = sum(0.5 d^2 if |d| < 1 else
|d| - 0.5); ClassNLLCriterion sizeAverage = true (mean over the batch rows).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import model as M
from .anchors import Anchors
from .detector import roi_crop_index
from .localizer import Localizer, trunk_layer_info
from .rect import Rect


def clean_anchors(examples, dims):
    """cleanAnchors (objective.lua:32-43): drop examples whose index lies outside the actual feature map."""
    return [e for e in examples if e[0].index[1] <= dims[e[0].layer - 1][1] and e[0].index[2] <= dims[e[0].layer - 1][2]]


def loss_and_gradient_image(desc, cfg, params, img, positives, negatives, dropout_masks=None, cnet_masks=None,
                            quant=None, act_quant=None, tail_quant="same", cnet_quant=None, inject_blocks=None):
    """One iteration of the per-image loop of lossAndGradient (objective.lua:65-198).
    positives: list of (anchor Rect with .layer/.aspect/.index, roi dict{rect, class_index}); negatives: list of
    (anchor,).  Returns (losses dict, grads dict name -> tensor, intermediates)."""
    p = {k: v.clone().requires_grad_(not k.endswith(("bn_mean", "bn_var"))) for k, v in params.items()}
    kh, kw = cfg["roi_pooling"]["kh"], cfg["roi_pooling"]["kw"]
    bgclass = cfg["class_count"] + 1
    localizer = Localizer(trunk_layer_info(desc["layers"], len(desc["layers"])))
    outputs = M.pnet_forward(desc, p, img, train=True, dropout_masks=dropout_masks, quant=quant, act_quant=act_quant,
                             tail_quant=tail_quant, inject_blocks=inject_blocks)
    dims = [tuple(o.shape) for o in outputs]
    positives = clean_anchors(positives, dims)
    negatives = clean_anchors(negatives, dims)
    fmap = outputs[4]
    C, fh, fw = fmap.shape
    cls_loss = torch.zeros(())
    reg_loss = torch.zeros(())
    pooled, cctarget, crtarget = [], [], []
    for anchor, roi in positives:  # objective.lua:91-120
        (c0, c1), y, x = anchor.index
        v = outputs[anchor.layer - 1][c0 - 1:c1, y - 1, x - 1]
        cls_loss = cls_loss + F.cross_entropy(v[0:2].unsqueeze(0), torch.tensor([0]), reduction="sum")
        reg_out = v[2:6]
        reg_target = torch.from_numpy(Anchors.inputToAnchor(anchor, roi["rect"]))
        reg_loss = reg_loss + 10 * F.smooth_l1_loss(reg_out, reg_target, reduction="sum", beta=1.0)
        reg_proposal = Anchors.anchorToInput(anchor, reg_out.detach().numpy())
        idx = roi_crop_index(roi["rect"], localizer, fh, fw)
        y0, y1, x0, x1 = idx
        pooled.append(F.adaptive_max_pool2d(fmap[:, y0:y1, x0:x1].unsqueeze(0), (kh, kw)).reshape(-1))
        cctarget.append(roi["class_index"])
        crtarget.append(torch.from_numpy(Anchors.inputToAnchor(reg_proposal, roi["rect"])))
    for (anchor,) in negatives:  # objective.lua:123-140
        (c0, c1), y, x = anchor.index
        v = outputs[anchor.layer - 1][c0 - 1:c1, y - 1, x - 1]
        cls_loss = cls_loss + F.cross_entropy(v[0:2].unsqueeze(0), torch.tensor([1]), reduction="sum")
        y0, y1, x0, x1 = roi_crop_index(anchor, localizer, fh, fw)
        pooled.append(F.adaptive_max_pool2d(fmap[:, y0:y1, x0:x1].unsqueeze(0), (kh, kw)).reshape(-1))
        cctarget.append(bgclass)
        crtarget.append(torch.zeros(4))
    total = cls_loss + reg_loss
    creg_loss = torch.zeros(())
    ccls_loss = torch.zeros(())
    inter = dict(outputs=[o.detach() for o in outputs], n_pos=len(positives), n_neg=len(negatives))
    if pooled:  # objective.lua:146-186
        cinput = torch.stack(pooled)
        crt = torch.stack(crtarget)
        crout, ccout = M.cnet_forward(desc, p, cinput, train=True, dropout_masks=cnet_masks, quant=cnet_quant, quant_heads=None)
        npos = len(positives)
        keep = torch.zeros_like(crout)
        keep[:npos] = 1.0
        crout = crout * keep  # crout[{{#p + 1, #roi_pool_state}, {}}]:zero()
        creg_loss = 10 * F.smooth_l1_loss(crout, crt, reduction="sum", beta=1.0)
        ccls_loss = F.nll_loss(ccout, torch.tensor(cctarget) - 1, reduction="mean")
        total = total + creg_loss + ccls_loss
        inter.update(cinput=cinput.detach(), crout=crout.detach(), ccout=ccout.detach(), crtarget=crt)
    total.backward()
    grads = {k: v.grad for k, v in p.items() if v.grad is not None}
    losses = dict(cls=cls_loss.item(), reg=reg_loss.item(), creg=creg_loss.item(), ccls=ccls_loss.item())
    return losses, grads, inter


def synthetic_examples(anchors, dims, img_w, img_h, n_pos, n_neg, n_gt, class_count, seed=0):
    """`n_gt` ground-truth boxes, positives / negatives drawn uniformly from the valid anchor
    indices (seeded); every positive is paired with the ground-truth box nearest to its anchor centre."""
    rng = np.random.default_rng(seed)
    gts = []
    for _ in range(n_gt):
        w, h = rng.uniform(40, 0.5 * img_w), rng.uniform(40, 0.5 * img_h)
        x, y = rng.uniform(0, img_w - w), rng.uniform(0, img_h - h)
        gts.append(dict(rect=Rect(x, y, x + w, y + h), class_index=int(rng.integers(1, class_count + 1))))

    def draw():
        layer = int(rng.integers(1, 5))
        _, hh, ww = dims[layer - 1]
        return anchors.get(layer, int(rng.integers(1, 4)), int(rng.integers(1, hh + 1)), int(rng.integers(1, ww + 1)))

    pos, neg = [], []
    for _ in range(n_pos):
        a = draw()
        cx, cy = a.center()
        roi = min(gts, key=lambda g: (g["rect"].center()[0] - cx) ** 2 + (g["rect"].center()[1] - cy) ** 2)
        pos.append((a, roi))
    for _ in range(n_neg):
        neg.append((draw(),))
    return pos, neg, gts
