"""Anchor labelling: Anchors:findPositive / sampleNegative (Anchors.lua:147-235) on the GPU against the
oracle's restatement of the Lua loops -- bit-exact index lists in the reference's order."""
import numpy as np
import pytest

from oracle import anchors as OA, model as OM
from oracle.rect import Rect


def _oracle_anchors():
    return OA.Anchors(OM.VGG_SMALL["layers"], OM.VGG_SMALL["anchor_nets"], OM.CFG_DUPLO["scales"])


def _rois(seed, n, w=800, h=450):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        bw, bh = rng.uniform(20, 300), rng.uniform(20, 250)
        x, y = rng.uniform(0, w - bw), rng.uniform(0, h - bh)
        out.append({"rect": Rect(x, y, x + bw, y + bh), "class_index": int(rng.integers(1, 17))})
    return out


def _ref(a):
    return (a.layer, a.aspect, a.index[1], a.index[2])


def test_oracle_sample_negative_rules():
    """The loop stops at `count` accepted anchors or after 500 consecutive rejections; three random values per trial."""
    oa = _oracle_anchors()
    img = Rect(0, 0, 800, 450)
    rng = np.random.default_rng(0)
    rnd = rng.integers(0, 2 ** 32, 3 * 4000, dtype=np.uint64)
    neg, trials = oa.sampleNegative(img, _rois(1, 4), 0.3, 64, rnd)
    assert len(neg) == 64 and trials >= 64
    for (a,) in neg:
        assert a.minX >= 0 and a.minY >= 0 and a.maxX <= 800 and a.maxY <= 450     # anchors inside the image (clip = image)
        assert all(Rect.IoU(r["rect"], a) <= 0.3 for r in _rois(1, 4))
    # a ROI covering everything with threshold -1 rejects every trial: 500 consecutive rejections end the loop
    neg, trials = oa.sampleNegative(img, [{"rect": Rect(0, 0, 800, 450)}], -1.0, 10, rnd)
    assert neg == [] and trials == 500


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,best,clip", [(0, 1, True, True), (1, 8, True, True), (2, 8, False, True), (3, 5, True, False),
                                              (4, 16, True, True)])
def test_gpu_find_positive_bit_exact(F, small_model, seed, n, best, clip):
    oa = _oracle_anchors()
    ga = F.Anchors(small_model)
    rois = _rois(seed, n)
    if seed == 4:   # tiny and huge boxes: no positive anchor -> the best-set branch; and boxes partly outside the image
        rois[0]["rect"] = Rect(100, 100, 112, 109)
        rois[1]["rect"] = Rect(-50, -40, 700, 500)
        rois[2]["rect"] = Rect(790, 440, 830, 470)
    img = Rect(0, 0, 800, 450) if clip else None
    want = oa.findPositive(rois, img, 0.6, 0.3, best)
    got = ga.findPositive(rois, img, 0.6, 0.3, best)
    assert [(_ref(a), id(r)) for a, r in got] == [(_ref(a), id(r)) for a, r in want]
    assert len(want) > 0
    for (a, _), (b, _) in zip(got, want):
        assert a.unpack() == b.unpack()


@pytest.mark.gpu
def test_gpu_find_positive_thresholds(F, small_model):
    """Low thresholds: thousands of positives per ROI (several 256-candidate chunks, ordered compaction across them)."""
    oa = _oracle_anchors()
    ga = F.Anchors(small_model)
    rois = _rois(7, 3)
    want = oa.findPositive(rois, Rect(0, 0, 800, 450), 0.05, 0.01, True)
    got = ga.findPositive(rois, Rect(0, 0, 800, 450), 0.05, 0.01, True)
    assert len(want) > 1000
    assert [_ref(a) for a, _ in got] == [_ref(a) for a, _ in want]
    assert ga.findPositive([], None, 0.6, 0.3, True) == []


@pytest.mark.gpu
@pytest.mark.parametrize("seed,count,thr", [(0, 128, 0.3), (1, 16, 0.3), (2, 256, 0.05), (3, 700, 0.3)])
def test_gpu_sample_negative_bit_exact(F, small_model, seed, count, thr):
    oa = _oracle_anchors()
    ga = F.Anchors(small_model)
    rois = _rois(seed + 10, 6)
    rng = np.random.default_rng(seed)
    rnd = rng.integers(0, 2 ** 32, 3 * 3000, dtype=np.uint64).astype(np.uint32)
    img = Rect(0, 0, 800, 450)
    want, trials = oa.sampleNegative(img, rois, thr, count, rnd)
    got, used, finished = ga.sampleNegative(img, rois, thr, count, rnd)
    assert [_ref(a) for (a,) in got] == [_ref(a) for (a,) in want]
    assert used == trials and finished and len(got) == count


@pytest.mark.gpu
def test_gpu_sample_negative_stopping_rules(F, small_model):
    ga = F.Anchors(small_model)
    oa = _oracle_anchors()
    rng = np.random.default_rng(5)
    rnd = rng.integers(0, 2 ** 32, 3 * 2000, dtype=np.uint64).astype(np.uint32)
    img = Rect(0, 0, 800, 450)
    # every trial rejected: the loop gives up after 500 consecutive rejections
    got, used, finished = ga.sampleNegative(img, [{"rect": Rect(0, 0, 800, 450)}], -1.0, 10, rnd)
    assert got == [] and used == 500 and finished
    # stream shorter than needed: reports that the rule did not fire
    got, used, finished = ga.sampleNegative(img, _rois(3, 2), 0.3, 500, rnd[:3 * 100])
    want, trials = oa.sampleNegative(img, _rois(3, 2), 0.3, 500, rnd[:3 * 100])
    assert used == 100 == trials and not finished
    assert [_ref(a) for (a,) in got] == [_ref(a) for (a,) in want]
    # mostly rejected with occasional accepts: the retry counter resets on every accept
    rois = [{"rect": Rect(0, 0, 800, 450)}]
    want, trials = oa.sampleNegative(img, rois, 0.02, 40, rnd)
    got, used, finished = ga.sampleNegative(img, rois, 0.02, 40, rnd)
    assert [_ref(a) for (a,) in got] == [_ref(a) for (a,) in want] and used == trials


@pytest.mark.gpu
def test_gpu_sample_negative_continued_stream(F, small_model):
    """A loop whose random stream is handed over in pieces (remaining count + carried run of rejections) returns what
    the uninterrupted loop returns: the accepted anchors, the number of trials, and the stopping rule."""
    ga = F.Anchors(small_model)
    oa = _oracle_anchors()
    rng = np.random.default_rng(9)
    rnd = rng.integers(0, 2 ** 32, 3 * 1500, dtype=np.uint64).astype(np.uint32)
    img = Rect(0, 0, 800, 450)
    rois = [{"rect": Rect(0, 0, 800, 450)}]          # threshold 0.02: most trials are rejected
    want, trials = oa.sampleNegative(img, rois, 0.02, 40, rnd)
    got, pos, need, retry, finished = [], 0, 40, 0, False
    for piece in (70, 130, 300, 1000):
        part, used, finished, retry = ga.sampleNegative(img, rois, 0.02, need, rnd[3 * pos:3 * (pos + piece)], retry=retry, return_retry=True)
        got += part
        pos += used
        need -= len(part)
        if finished:
            break
    assert finished and pos == trials
    assert [_ref(a) for (a,) in got] == [_ref(a) for (a,) in want]
    # the give-up rule across pieces: 500 consecutive rejections in total
    part, used, finished, retry = ga.sampleNegative(img, rois, -1.0, 5, rnd[:3 * 300], return_retry=True)
    assert part == [] and used == 300 and not finished and retry == 300
    part, used, finished, retry = ga.sampleNegative(img, rois, -1.0, 5, rnd[3 * 300:3 * 900], retry=retry, return_retry=True)
    assert part == [] and used == 200 and finished


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,thr", [(0, 4, 0.3), (1, 8, 0.25), (2, 8, 0.9), (3, 2, 0.0)])
def test_gpu_nearby_negative_bit_exact(F, small_model, seed, n, thr):
    """cfg.nearby_aversion (BatchIterator.lua:206-217): for every positive anchor, Anchors:findNearby(center) filtered by
    Rect.IoU(p, a) < negative_threshold -- the device list equals the restated Lua loops entry for entry, in order."""
    oa = _oracle_anchors()
    ga = F.Anchors(small_model)
    positive = oa.findPositive(_rois(seed, n), Rect(0, 0, 800, 450), 0.6, 0.3, True)
    assert len(positive) > 0
    want = []
    for p in positive:
        cx, cy = p[0].center()
        for a in oa.findNearby(cx, cy):
            if Rect.IoU(p[0], a) < thr:
                want.append(a)
    got = ga.nearbyNegative(positive, thr)
    assert [_ref(a) for (a,) in got] == [_ref(a) for a in want]
    if thr >= 0.25:
        assert len(want) > 0
    # the host mirror's own findNearby agrees with the oracle's bins
    for p in positive[:20]:
        cx, cy = p[0].center()
        assert [_ref(a) for a in ga.findNearby(cx, cy)] == [_ref(a) for a in oa.findNearby(cx, cy)]
