"""CPU tests of the oracle's restatement of lossAndGradient (objective.lua:45-218) on a small frame: the criteria's
hand-computable values and the structural facts the reference's code implies."""
import math

import numpy as np
import torch

from oracle import anchors as OA, model as OM, objective as OO
from oracle.rect import Rect


def _setup(h=122, w=192, seed=0):
    desc, cfg = OM.VGG_SMALL, OM.CFG_DUPLO
    p = OM.init_params(desc, cfg, seed=seed, randomize_aux=True)
    anchors = OA.Anchors(desc["layers"], desc["anchor_nets"], cfg["scales"])
    img = OM.synthetic_frame(h, w, seed=seed)
    with torch.no_grad():
        dims = [tuple(o.shape) for o in OM.pnet_forward(desc, p, img)]
    return desc, cfg, p, anchors, img, dims


def test_losses_match_hand_computation():
    desc, cfg, p, anchors, img, dims = _setup()
    pos, neg, _ = OO.synthetic_examples(anchors, dims, 192, 122, 6, 5, 2, cfg["class_count"], seed=1)
    g = torch.Generator().manual_seed(0)
    dm = {"b%d_c1" % (i + 1): torch.ones(l["filters"]) for i, l in enumerate(desc["layers"]) if l["dropout"]}
    cm = {"fc1": torch.ones(1024), "fc2": torch.ones(512)}
    losses, grads, inter = OO.loss_and_gradient_image(desc, cfg, p, img, pos, neg, dropout_masks=dm, cnet_masks=cm)
    outs = inter["outputs"]
    # proposal classification: -log softmax(v[1:2])[target], summed (CrossEntropyCriterion on one sample)
    want_cls, want_reg = 0.0, 0.0
    for a, roi in pos:
        (c0, c1), y, x = a.index
        v = outs[a.layer - 1][c0 - 1:c1, y - 1, x - 1].double().numpy()
        want_cls += -(v[0] - np.logaddexp(v[0], v[1]))
        t = OA.Anchors.inputToAnchor(a, roi["rect"]).astype(np.float64)
        d = np.abs(v[2:6] - t)
        want_reg += 10 * np.sum(np.where(d < 1, 0.5 * d * d, d - 0.5))
    for (a,) in neg:
        (c0, c1), y, x = a.index
        v = outs[a.layer - 1][c0 - 1:c1, y - 1, x - 1].double().numpy()
        want_cls += -(v[1] - np.logaddexp(v[0], v[1]))
    assert math.isclose(losses["cls"], want_cls, rel_tol=1e-4)
    assert math.isclose(losses["reg"], want_reg, rel_tol=1e-4)
    # detection stage: negatives' bbox outputs are zeroed before the loss (objective.lua:170); NLL is a mean
    assert torch.all(inter["crout"][len(pos):] == 0)
    cc = inter["ccout"]
    tgt = [roi["class_index"] for _, roi in pos] + [cfg["class_count"] + 1] * len(neg)
    assert math.isclose(losses["ccls"], -float(sum(cc[i, t - 1] for i, t in enumerate(tgt))) / len(tgt), rel_tol=1e-5)
    # every learnable tensor receives a gradient; BN running statistics do not
    learnable = [k for k in p if not k.endswith(("bn_mean", "bn_var"))]
    assert set(grads) == set(learnable)
    assert all(torch.isfinite(v).all() for v in grads.values())


def test_negatives_only_have_no_regression_gradient():
    desc, cfg, p, anchors, img, dims = _setup(seed=2)
    _, neg, _ = OO.synthetic_examples(anchors, dims, 192, 122, 0, 8, 1, cfg["class_count"], seed=3)
    dm = {"b%d_c1" % (i + 1): torch.ones(l["filters"]) for i, l in enumerate(desc["layers"]) if l["dropout"]}
    cm = {"fc1": torch.ones(1024), "fc2": torch.ones(512)}
    losses, grads, _ = OO.loss_and_gradient_image(desc, cfg, p, img, [], neg, dropout_masks=dm, cnet_masks=cm)
    assert losses["reg"] == 0.0 and losses["creg"] == 0.0
    assert float(grads["reg.weight"].abs().max()) == 0.0 and float(grads["reg.bias"].abs().max()) == 0.0
    # the 1x1 head convs only receive gradient on the two class rows of the sampled aspects
    for hname in ("h1", "h2", "h3", "h4"):
        if hname + "_out.weight" not in grads:  # no example landed on this head
            continue
        gw = grads[hname + "_out.weight"].reshape(18, -1)
        assert float(gw[[2, 3, 4, 5, 8, 9, 10, 11, 14, 15, 16, 17]].abs().max()) == 0.0


def test_clean_anchors_drops_out_of_range_indices():
    desc, cfg, p, anchors, img, dims = _setup()
    a_ok = anchors.get(1, 1, 1, 1)
    a_bad = anchors.get(4, 2, dims[3][1] + 1, 1)  # one row below head 4's map
    kept = OO.clean_anchors([(a_ok,), (a_bad,)], dims)
    assert len(kept) == 1 and kept[0][0] is a_ok
