"""Frame normalisation before pnet:forward: oracle restatement and GPU parity (tolerance 1e-5 of the
value range: the un-vendored TH convolution's accumulation order is unknown, statistics are reduced in double)."""
import numpy as np
import pytest
import torch

from oracle import preprocess as OP


def test_oracle_gaussian_and_border_coefficient():
    g = OP.gaussian1D(7)
    assert g.shape == (7,) and g[3] == pytest.approx(1.0) and g[0] == pytest.approx(np.exp(-((3 / 1.75) ** 2) / 2), rel=1e-6)
    assert np.allclose(g, g[::-1])
    # a constant plane: the subtractive stage removes it exactly everywhere (the coefficient map undoes the zero padding)
    x = torch.full((20, 31), 3.0)
    k = (g / g.sum()).astype(np.float32)
    coef = OP._mean_estimator(torch.ones_like(x), k)
    assert coef[10, 15] == pytest.approx(1.0, rel=1e-6) and coef[0, 0] < 0.6
    s = x - OP._mean_estimator(x, k) / coef
    assert s.abs().max().item() < 1e-5


def test_oracle_normalize_statistics():
    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.uniform(0, 1, (3, 45, 80)).astype(np.float32))
    out = OP.normalize_frame(img, contrastive_width=0)
    for c in range(3):
        assert abs(out[c].mean().item()) < 1e-6 and out[c].std().item() == pytest.approx(1.0, rel=1e-5)
    flat = torch.zeros(3, 8, 8)
    assert torch.equal(OP.normalize_frame(flat, contrastive_width=0), flat)   # std <= 1e-8: channel left alone


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", [(45, 80), (450, 800), (61, 97)])
@pytest.mark.parametrize("yuv", [False, True])
def test_gpu_normalize_frame(F, small_model, h, w, yuv):
    rng = np.random.default_rng(h + w)
    base = rng.uniform(0, 1, (3, h, w)).astype(np.float32)
    base[:, h // 3:h // 2, w // 4:w // 2] += 0.8          # structure, so local statistics vary
    img = torch.from_numpy(base)
    want = OP.normalize_frame(img, rgb_to_yuv=yuv)
    got = small_model.normalize_frame(img.clone().cuda(), rgb2yuv=yuv).cpu()
    scale = want.abs().max().item()
    assert (got - want).abs().max().item() <= 1e-5 * max(scale, 1.0) * 10
    # stages alone
    got2 = small_model.normalize_frame(img.clone().cuda(), rgb2yuv=yuv, contrastive_width=0).cpu()
    want2 = OP.normalize_frame(img, rgb_to_yuv=yuv, contrastive_width=0)
    assert (got2 - want2).abs().max().item() <= 2e-6 * max(want2.abs().max().item(), 1.0)


def test_find_target_size_host_equals_oracle():
    import frcnn_b200 as F
    for (w, h) in [(1280, 720), (720, 1280), (800, 450), (500, 375), (333, 500), (4000, 600), (600, 4000), (17, 17), (1, 1), (1000, 999)]:
        for tss, mps in [(450, 1000), (480, 1000), (600, 800)]:
            assert F.find_target_size(w, h, tss, mps) == OP.find_target_size(w, h, tss, mps)
    assert OP.find_target_size(1280, 720, 450, 1000) == (800, 450)        # the headline frame size
    assert OP.find_target_size(4000, 600, 450, 1000) == (1000, 150)       # max_pixel_size caps the long side


def test_oracle_scale_known_answers():
    # enlarging 3 -> 5 samples: scale (3-1)/(5-1) = 0.5 -> positions 0, .5, 1, 1.5 and the last sample copied
    a = np.array([[0.0, 2.0, 6.0]], dtype=np.float32)
    assert np.array_equal(OP._scale_linear_1d(a, 5), np.array([[0.0, 1.0, 2.0, 4.0, 6.0]], dtype=np.float32))
    # shrinking 6 -> 4: intervals of 1.5 samples with fractional end weights
    b = np.arange(6, dtype=np.float32)[None]
    want = [(0 + 0.5 * 1) / 1.5, (0.5 * 1 + 2) / 1.5, (3 + 0.5 * 4) / 1.5, (0.5 * 4 + 5) / 1.5]
    assert np.allclose(OP._scale_linear_1d(b, 4), np.array([want], dtype=np.float32), rtol=1e-6)
    c = np.random.default_rng(0).uniform(0, 1, (3, 9, 7)).astype(np.float32)
    assert np.array_equal(OP.scale_image(c, 7, 9), c)                      # same size: a copy
    assert OP.scale_image(c, 12, 4).shape == (3, 4, 12)
    assert abs(float(OP.scale_image(np.full((1, 20, 30), 0.25, np.float32), 11, 47).max()) - 0.25) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("sh,sw,dh,dw", [(720, 1280, 450, 800), (375, 500, 450, 600), (100, 60, 37, 91), (5, 1, 9, 4), (64, 64, 64, 64)])
def test_gpu_scale_frame_bit_exact_vs_restatement(F, small_model, sh, sw, dh, dw):
    rng = np.random.default_rng(sh * 7 + dw)
    img = rng.uniform(0, 1, (3, sh, sw)).astype(np.float32)
    want = OP.scale_image(img, dw, dh)
    got = small_model.scale_frame(torch.from_numpy(img).cuda(), dw, dh).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_frame_prefetcher_pipeline(F, small_model):
    """The asynchronous upload + resize + normalise pipeline (side stream, events) yields, in submission order, exactly what
    the same library calls yield synchronously; a consumer on another stream sees complete frames."""
    rng = np.random.default_rng(5)
    frames = [torch.from_numpy(rng.uniform(0, 1, (3, 360 + 40 * i, 640)).astype(np.float32)).pin_memory() for i in range(5)]
    norm = dict(rgb2yuv=True, centering=True, scaling=True, contrastive_width=7)
    pf = F.FramePrefetcher(device=0, depth=2, target_smaller_side=450, max_pixel_size=1000, normalization=norm)
    try:
        want = []
        for f in frames:
            w, h = F.find_target_size(f.shape[2], f.shape[1], 450, 1000)
            x = small_model.scale_frame(f.cuda(), w, h)
            want.append(small_model.normalize_frame(x, rgb2yuv=True).clone())
        got = []
        pf.submit(frames[0])
        pf.submit(frames[1])
        for i in range(len(frames)):
            x = pf.get()
            got.append((x.sum() * 0 + x).clone())          # consumed on the current stream, after the event
            if i + 2 < len(frames):
                pf.submit(frames[i + 2])
        torch.cuda.synchronize()
        for g, wnt in zip(got, want):
            assert g.shape == wnt.shape and torch.equal(g, wnt)
        assert pf.pending() == 0
    finally:
        pf.close()
