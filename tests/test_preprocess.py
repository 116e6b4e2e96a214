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
