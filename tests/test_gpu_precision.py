"""Precision contract of the tensor-core path against the PURE fp32 restatement of the reference, end to end from the
frame at the headline configurations (BASELINE configs[1]: vgg_small 800x450; configs[3]: vgg_large 1000x600).

The reference runs pnet / cnet in fp32 (cunn).  The CUDA path rounds conv / Linear operands to 16 bits and
accumulates in fp32, so the continuous outputs differ and anchors sitting on the p > 0.95 edge (Detector.lua:54) can
land on the other side.  oracle/compare.py measures exactly that; this test states the bars and writes the measured
report to gpurun_out/precision_<model>.json (copied to profiles/ by hand).

Detections are reproducible run to run (deterministic split-K slices everywhere on the detect path), asserted bit
for bit below."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import compare as OC, detector as OD, model as OM

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Stated bars per operand format.  Continuous outputs per element vs the fp32 path: |cuda - fp32| <= atol + rtol * |fp32|.
# Discrete stages: Jaccard index of the match / candidate / winner sets of the two paths.  The greedy NMS of the
# reference orders boxes by their BOTTOM EDGE (nms.lua:41-42), so sub-pixel differences of the regressed
# boxes re-order near-ties and the suppression chain amplifies them: ONE reordered near-tie swaps which anchor of a
# cluster survives, and everything it would have suppressed.  Two fp16 kernels that differ only in their fp32 summation
# order (both within 3e-3 of the fp32 maps) were measured at 1.00 and 0.78 candidate Jaccard on the same vgg_large
# frame, so the set bars are stated on the MEAN over several frames, next to a geometric bar that does not depend on
# which anchor of a cluster won: the share of the fp32 winners covered by a CUDA winner of the same class at
# IoU >= 0.7 (and vice versa).  With bf16 operands (8 significand bits, boxes move by up to ~2 px) the candidate sets
# drift apart even though 99 % of the matches agree; fp16 operands (11 bits) are the evaluate-mode default.
# Measured values: profiles/r2_precision_*.json.
BARS = {
    "fp16": dict(head=(6e-3, 4e-3), feat=(2e-3, 4e-3), matches=0.99, candidates=0.75, winners=0.85, covered=0.90, r2=0.005),
    "bf16": dict(head=(4e-2, 2e-2), feat=(1e-2, 2e-2), matches=0.98, candidates=0.30, winners=0.20, covered=0.50, r2=0.03),
}
FRAME_SEEDS = {"small": (2, 3, 4, 5), "large": (2, 3, 4)}


def _model(F, which):
    if which == "small":
        m = F.vgg_small(F.duplo_cfg)
        desc, cfg = OM.VGG_SMALL, OM.CFG_DUPLO
    else:
        m = F.vgg_large(F.imgnet_cfg)
        desc, cfg = OM.VGG_LARGE, OM.CFG_IMAGENET
    p = OM.detecting_params(OM.init_params(desc, cfg, seed=0, randomize_aux=True))
    m.load_params(p)
    return m, desc, cfg, p


@pytest.mark.parametrize("which,h,w", [("small", 450, 800), ("large", 600, 1000)])
def test_precision_vs_fp32_reference_path(F, which, h, w):
    m, desc, cfg, p = _model(F, which)
    try:
        out_dir = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        reports = {"fp16": [], "bf16": []}
        oracle = OD.Detector(desc, cfg, p)
        for seed in FRAME_SEEDS[which]:
            img = OM.synthetic_frame(h, w, seed=seed)
            fp32 = oracle.detect(img, return_intermediates=True)   # the reference's arithmetic, from the frame
            want = fp32[1]["outputs"]
            assert len(fp32[1]["matches"]) > 100, "the test weights must put a realistic number of anchors above 0.95"
            for precision in ("fp16", "bf16"):
                m.set_eval_precision(precision)
                det = F.Detector(m)
                winners = det.detect(img.numpy())
                stats = det.stats()
                maps = [o.cpu() for o in m.pnet.forward(img.cuda())]
                rep = OC.precision_report(desc, cfg, p, img, maps, winners, fp32_result=fp32,
                                          quant=OM.fp16_round if precision == "fp16" else OM.bf16_round)
                rep["config"] = dict(model="vgg_" + which, h=h, w=w, frame_seed=seed, operands=precision,
                                     weights="init_params(seed=0, randomize_aux) + detecting_params")
                rep["cuda_stats"] = stats
                reports[precision].append(rep)
                print(json.dumps({k: v for k, v in rep.items() if k != "maps"}))
                bars = BARS[precision]
                # the discrete stages computed from the CUDA maps by the oracle are what the kernels produced
                assert stats["matches"] == rep["matches"]["cuda"] and stats["candidates"] == rep["candidates"]["cuda"]
                for i, (g, wnt) in enumerate(zip(maps, want)):
                    atol, rtol = bars["head"] if i < 4 else bars["feat"]
                    err = (g - wnt).abs()
                    assert bool((err <= atol + rtol * wnt.abs()).all()), (precision, seed, i, float(err.max()))
                assert rep["matches"]["jaccard"] >= bars["matches"], (precision, seed, rep["matches"])
                assert rep["winners"]["r2_max_rel_to_box"] <= bars["r2"], (precision, seed, rep["winners"])
                for mrec in rep["maps"][:4]:
                    # the anchors that flipped all sat within the measured error of the threshold
                    assert mrec["flips"] <= mrec["anchors_within_err_of_threshold"]
        for precision, reps in reports.items():
            bars = BARS[precision]
            mean = lambda f: float(np.mean([f(r) for r in reps]))  # noqa: E731
            summary = dict(frames=len(reps), operands=precision, model="vgg_" + which, h=h, w=w,
                           matches_jaccard_min=min(r["matches"]["jaccard"] for r in reps),
                           candidates_jaccard_mean=mean(lambda r: r["candidates"]["jaccard"]),
                           winners_jaccard_mean=mean(lambda r: r["winners"]["jaccard"]),
                           fp32_winners_covered_mean=mean(lambda r: r["winners"]["fp32_covered_by_cuda"]),
                           cuda_winners_covered_mean=mean(lambda r: r["winners"]["cuda_covered_by_fp32"]),
                           head_max_abs=max(x["max_abs"] for r in reps for x in r["maps"][:4]),
                           feature_max_abs=max(r["maps"][4]["max_abs"] for r in reps),
                           r2_max_rel_to_box=max(r["winners"]["r2_max_rel_to_box"] for r in reps),
                           bars=bars)
            with open(os.path.join(out_dir, "precision_%s_%s.json" % (which, precision)), "w") as f:
                json.dump(dict(summary=summary, frames=reps), f, indent=1)
            print(json.dumps(summary))
            assert summary["candidates_jaccard_mean"] >= bars["candidates"], summary
            assert summary["winners_jaccard_mean"] >= bars["winners"], summary
            assert summary["fp32_winners_covered_mean"] >= bars["covered"], summary
            assert summary["cuda_winners_covered_mean"] >= bars["covered"], summary
    finally:
        m.close()


def test_detect_is_bit_reproducible(F):
    """Same frame, same context, eager and graph replay, latency and throughput schedules run twice each: identical
    winners down to the last bit of every double / float (no atomics or reduce-adds on the detect path)."""
    m, desc, cfg, p = _model(F, "small")
    try:
        img = OM.synthetic_frame(225, 400, seed=5).cuda()
        det = F.Detector(m)

        def snapshot():
            return [(x["class"], x["l"], x["a"].aspect, x["a"].index[1], x["a"].index[2], x["r2"].unpack(),
                     float(x["confidence"]).hex(), float(x["p"]).hex()) for x in det.detect(img)]

        for sched in ("latency", "throughput"):
            m.set_schedule(sched)
            runs = [snapshot() for _ in range(5)]   # eager, capture, replays
            assert len(runs[0]) > 0
            for r in runs[1:]:
                assert r == runs[0]
        reg1, cls1 = None, None
        x = torch.randn(200, 6 * 6 * 384, device="cuda")
        for _ in range(3):
            reg, cls = m.cnet.forward(x)
            if reg1 is None:
                reg1, cls1 = reg.clone(), cls.clone()
            assert torch.equal(reg, reg1) and torch.equal(cls, cls1)
    finally:
        m.close()


def test_detect_grows_candidate_capacity(F, small_model):
    """The reference's match list is unbounded (Detector.lua:59).  A frame with more matches than the candidate buffers
    hold makes the library grow them and run the frame again -- no FRCNN_E_OVERFLOW -- up to every anchor passing
    (26 544 at 800x450, past the 8192-box CTA-level NMS: the radix-sort path)."""
    from oracle import nms as ONMS
    p = dict(small_model.oracle_params)
    for shift, expect_all in ((3.0, False), (30.0, True)):
        q = {k: v.clone() for k, v in p.items()}
        for k in q:
            if k.endswith("_out.bias"):
                q[k][0::6] += shift
        m = F.vgg_small(F.duplo_cfg)
        m.load_params(q)
        try:
            det = F.Detector(m)
            img = OM.synthetic_frame(450, 800, seed=1)
            det.detect(img.numpy())
            st = det.stats()
            maps = [o.cpu() for o in m.pnet.forward(img.cuda())]
            od = OD.Detector(OM.VGG_SMALL, OM.CFG_DUPLO, q, quant=OM.fp16_round, quant_heads=None)
            matches = OD.decode(maps, od.anchors, OD.Rect(0, 0, 800, 450))
            assert st["matches"] == len(matches)
            assert st["matches"] > 4096
            if expect_all:
                assert st["matches"] > 8192
            bb = np.stack([x["r"].totensor() for x in matches])
            assert st["candidates"] == len(ONMS.nms(bb, 0.25, None))
            # the context keeps working at the grown capacity (and replays its graph again where it can)
            for _ in range(3):
                det.detect(img.numpy())
                assert det.stats()["matches"] == st["matches"] and det.stats()["candidates"] == st["candidates"]
        finally:
            m.close()
