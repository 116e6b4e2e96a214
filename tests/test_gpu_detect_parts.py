"""GPU parity tests of the detector-side kernels through the C ABI: RPN decode (Detector.lua:36-66), ROI pooling
(objective.lua:5-13 + nn.SpatialAdaptiveMaxPooling), cnet forward, against the oracle.

Bars: anchor enumeration / candidate order / argmax indices bit-exact; fp32 foreground log-probabilities within
1 ulp; decoded boxes within 1e-12 relative (double math; device exp() vs libm exp() may differ in the last ulp);
ROI max-pool values bit-exact (max is exact); cnet within the stated bf16-operand tolerance."""
import os

import numpy as np
import pytest
import torch

from oracle import anchors as OA, detector as OD, localizer as OL, model as OM
from oracle.rect import Rect as ORect

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _check_matches(got, want_anchor, want_logp, want_r, want_box):
    assert len(got) == len(want_anchor)
    ga = np.array([[m["l"], m["a"].aspect, m["a"].index[1], m["a"].index[2]] for m in got], np.int32).reshape(-1, 4)
    assert np.array_equal(ga, np.asarray(want_anchor).reshape(-1, 4))  # bit-exact enumeration + order
    if len(got) == 0:
        return
    lp = np.array([m["p"] for m in got], np.float32)
    ulp = np.abs(lp.view(np.int32).astype(np.int64) - np.asarray(want_logp, np.float32).view(np.int32).astype(np.int64))
    assert ulp.max() <= 1
    r = np.array([m["r"].unpack() for m in got])
    assert np.allclose(r, want_r, rtol=1e-12, atol=1e-9)
    box = np.stack([m["box"] for m in got])
    assert np.allclose(box, want_box, rtol=2e-7, atol=1e-5)


def test_decode_golden(F, small_model):
    g = np.load(os.path.join(GOLD, "decode.npz"))
    heads = [torch.from_numpy(g["head%d" % i]).cuda() for i in range(4)]
    det = F.Detector(small_model)
    got = det.decode(heads, 122, 192)
    assert len(got) > 50
    _check_matches(got, g["anchor"], g["logp"], g["r"], g["box"])


@pytest.mark.parametrize("h,w,shift", [(450, 800, 0.0), (450, 800, 2.5), (122, 192, 6.0), (300, 333, 1.0)])
def test_decode_vs_oracle(F, small_model, h, w, shift):
    """Full-size maps, none / few / nearly-all anchors passing the 0.95 threshold, odd sizes."""
    dims = small_model.output_dims(h, w)[:4]
    g = torch.Generator().manual_seed(h * 7 + w)
    heads = [torch.randn(d, generator=g) * 1.5 for d in dims]
    for t in heads:
        t[0::6] += shift
        t[4::6] *= 0.3
        t[5::6] *= 0.3
    oa = OA.Anchors(OM.VGG_SMALL["layers"], OM.VGG_SMALL["anchor_nets"], OM.CFG_DUPLO["scales"])
    want = OD.decode(heads, oa, ORect(0, 0, w, h))
    det = F.Detector(small_model)
    if len(want) > 4096:
        with pytest.raises(F.FrcnnError) as e:
            det.decode([t.cuda() for t in heads], h, w)
        assert e.value.code == 6  # FRCNN_E_OVERFLOW: more matches than the candidate capacity
        return
    got = det.decode([t.cuda() for t in heads], h, w)
    _check_matches(got, [[x["l"], x["a"].aspect, x["a"].index[1], x["a"].index[2]] for x in want],
                   [x["p"] for x in want], np.array([x["r"].unpack() for x in want]).reshape(-1, 4),
                   np.array([x["r"].totensor() for x in want]).reshape(-1, 4))


def test_decode_threshold_edge(F, small_model):
    """exp(c[1]) > 0.95 is evaluated in double on the fp32 log-probability (Detector.lua:54): logits chosen so the
    probability sits within a few ulp of the threshold on both sides."""
    dims = small_model.output_dims(122, 192)[:4]
    heads = [torch.zeros(d) for d in dims]
    base = float(np.log(0.95 / 0.05))
    offs = torch.linspace(-3e-6, 3e-6, dims[0][1] * dims[0][2]).reshape(dims[0][1], dims[0][2])
    heads[0][0] = base + offs
    oa = OA.Anchors(OM.VGG_SMALL["layers"], OM.VGG_SMALL["anchor_nets"], OM.CFG_DUPLO["scales"])
    want = OD.decode(heads, oa, ORect(0, 0, 192, 122))
    got = F.Detector(small_model).decode([t.cuda() for t in heads], 122, 192)
    assert 0 < len(want) < dims[0][1] * dims[0][2]
    assert [(m["l"], m["a"].aspect, m["a"].index[1], m["a"].index[2]) for m in got] == \
           [(x["l"], x["a"].aspect, x["a"].index[1], x["a"].index[2]) for x in want]


def test_roi_pool_vs_oracle(F, small_model):
    g = np.load(os.path.join(GOLD, "geometry_nms.npz"))
    gen = torch.Generator().manual_seed(3)
    fmap = torch.randn(384, 29, 50, generator=gen)
    loc = OL.Localizer(OL.trunk_layer_info(OM.VGG_SMALL["layers"], 4))
    rects, want, want_idx = [], [], []
    for r in g["roi_in"]:
        if OD.roi_crop_index(ORect(*r), loc, 29, 50) is None:
            continue
        o, ind, (y0, y1, x0, x1) = OD.roi_pool(fmap, ORect(*r), loc)
        # oracle indices are relative to the crop; the library reports flat indices into the feature plane
        ind = ind.reshape(384, 36)
        cw = x1 - x0
        ind = (ind // cw + y0) * 50 + (ind % cw + x0)
        rects.append(F.Rect(*r)); want.append(o); want_idx.append(ind.reshape(-1))
    assert len(rects) > 40
    out, arg = F.extract_roi_pooling_input(small_model, rects, fmap.cuda())
    assert torch.equal(out.cpu(), torch.stack(want))  # bit-exact: max of the same fp32 values
    assert torch.equal(arg.cpu().long(), torch.stack(want_idx))
    # tiny crops (smaller than 6x6) replicate cells; rects hanging over the border clip (Q8)
    small = [F.Rect(0, 0, 16, 16), F.Rect(790, 440, 800, 450), F.Rect(-40, -40, 10, 10), F.Rect(300, 200, 301, 201)]
    out, arg = F.extract_roi_pooling_input(small_model, small, fmap.cuda())
    for i, r in enumerate(small):
        o, _, _ = OD.roi_pool(fmap, ORect(*r.unpack()), loc)
        assert torch.equal(out[i].cpu(), o)


def test_roi_pool_empty_crop_raises(F, small_model):
    """A rect whose feature-space max clips to 0 makes the reference index row 0 and raise (objective.lua:11, Q8)."""
    fmap = torch.zeros(384, 29, 50).cuda()
    loc = OL.Localizer(OL.trunk_layer_info(OM.VGG_SMALL["layers"], 4))
    bad = ORect(-500, -500, -400, -400)
    assert OD.roi_crop_index(bad, loc, 29, 50) is None
    with pytest.raises(F.FrcnnError) as e:
        F.extract_roi_pooling_input(small_model, [F.Rect(*bad.unpack())], fmap)
    assert e.value.code == 5  # FRCNN_E_ROI_EMPTY


@pytest.mark.parametrize("R", [1, 7, 128, 300])
def test_cnet_forward(F, small_model, R):
    """cnet:forward (model_utilities.lua:76-108), evaluate mode.  Operands are rounded to fp16 for the tensor cores,
    accumulation is fp32 with a fixed summation order (deterministic split-K slices): compared per element with the pure
    fp32 oracle (the reference's arithmetic) and, twice as tight, with the oracle on fp16-rounded operands."""
    g = torch.Generator().manual_seed(R)
    x = torch.randn(R, 13824, generator=g).abs()  # ROI max-pool outputs are mostly positive
    p = small_model.oracle_params
    reg, cls = small_model.cnet.forward(x.cuda())
    with torch.no_grad():
        reg_q, cls_q = OM.cnet_forward(OM.VGG_SMALL, p, x, quant=OM.fp16_round, quant_heads=None)
        reg_f, cls_f = OM.cnet_forward(OM.VGG_SMALL, p, x)
    # the two final Linear layers run in fp32 on the GPU: only fc1/fc2 operands are quantised there
    assert torch.allclose(reg.cpu(), reg_q, rtol=2e-3, atol=2e-3)
    assert torch.allclose(cls.cpu(), cls_q, rtol=2e-3, atol=2e-3)
    assert torch.allclose(reg.cpu(), reg_f, rtol=4e-3, atol=4e-3)
    assert torch.allclose(cls.cpu(), cls_f, rtol=4e-3, atol=4e-3)
    assert torch.allclose(torch.logsumexp(cls, 1).cpu(), torch.zeros(R), atol=1e-5)
    # bf16 operands (FRCNN_PREC_BF16) for comparison: the previous, 8x looser contract
    small_model.set_eval_precision("bf16")
    try:
        reg_b, cls_b = small_model.cnet.forward(x.cuda())
    finally:
        small_model.set_eval_precision("fp16")
    assert torch.allclose(reg_b.cpu(), reg_f, rtol=5e-2, atol=5e-2)
    assert torch.allclose(cls_b.cpu(), cls_f, rtol=5e-2, atol=5e-2)


def test_amp_module_slot_on_strided_views(F, small_model):
    """The `amp` slot the way objective.lua:118-119,183-184 and Detector.lua:96-97 use it: extract_roi_pooling_input
    returns a NON-contiguous crop view, amp:forward pools it, amp.indices is cloned and put back, amp:backward returns
    gradInput of the view's shape.  Values bit-exact vs the oracle's per-ROI restatement and vs the batched
    frcnn_roi_pool_forward; winners = the first maximum in row-major order (ties included); backward = scatter-add at the
    winners (overlapping bins of crops smaller than 6 x 6 accumulate)."""
    rng = np.random.default_rng(11)
    C, H, W = 384, 29, 50
    fmap = torch.from_numpy(rng.integers(-3, 4, size=(C, H, W)).astype(np.float32))   # many ties
    loc = F.Localizer(small_model, small_model.n_heads + 1)
    oloc = OL.Localizer(OL.trunk_layer_info(OM.VGG_SMALL["layers"], 4))
    amp = F.SpatialAdaptiveMaxPooling(small_model, 6, 6)
    rects = [(100, 50, 300, 250), (0, 0, 800, 450), (790, 440, 800, 450), (0, 0, 16, 16), (333.3, 120.7, 390.2, 200.1), (5, 5, 40, 300)]
    g = fmap.cuda()
    batched, _ = F.extract_roi_pooling_input(small_model, [F.Rect(*r) for r in rects], g)
    for k, r in enumerate(rects):
        view, idx = F.roi_pooling_view(loc, F.Rect(*r), g)
        assert not view.is_contiguous() or view.shape[2] == W
        out = amp.forward(view)
        want, ind, (y0, y1, x0, x1) = OD.roi_pool(fmap, ORect(*r), oloc)
        assert idx[1] == (y0 + 1, y1) and idx[2] == (x0 + 1, x1)
        assert torch.equal(out.cpu().reshape(-1), want.reshape(-1))
        assert torch.equal(out.reshape(-1), batched[k])
        ref, ref_idx = torch.nn.functional.adaptive_max_pool2d(view.cpu().contiguous(), (6, 6), return_indices=True)
        assert torch.equal(out.cpu(), ref)
        assert torch.equal(amp.indices.cpu().to(torch.int64), ref_idx)          # first maximum in row-major order
        # objective.lua:119 / :183: indices cloned, other ROIs pooled in between, indices put back before backward
        saved = amp.indices.clone()
        amp.forward(F.roi_pooling_view(loc, F.Rect(10, 10, 200, 200), g)[0])
        amp.indices = saved
        dout = torch.from_numpy(rng.standard_normal((C, 6, 6)).astype(np.float32)).cuda()
        gi = amp.backward(view, dout)
        assert tuple(gi.shape) == tuple(view.shape)
        want_gi = torch.zeros(view.shape, dtype=torch.float64).reshape(C, -1)
        want_gi.scatter_add_(1, ref_idx.reshape(C, -1), dout.cpu().double().reshape(C, -1))
        assert torch.allclose(gi.cpu().double().reshape(C, -1), want_gi, atol=1e-5)
    with pytest.raises(F.FrcnnError):
        amp.forward(g[:, 3:3, 2:9])     # an empty crop: cunn raises as well
