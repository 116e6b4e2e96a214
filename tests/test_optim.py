"""The optimiser step around the path: oracle restatement of optim.rmsprop (CPU) and bit-exact
parity of the fused CUDA pass (frcnn_rmsprop_step) with it."""
import numpy as np
import pytest

from oracle import optim as OO


def test_oracle_rmsprop_known_answer():
    """Hand-evaluated first steps (fp32): x=1, g=0.5, lr=0.01, alpha=0.99: m1 = 0.01*0.25 = 0.0025, tmp = 0.05 + 1e-8,
    x1 = 1 - 0.01*0.5/0.05 = 0.9; second step with the same gradient: m2 = 0.0025*0.99 + 0.0025 = 0.004975."""
    x = np.array([1.0], np.float32)
    st = {}
    OO.rmsprop_step(x, np.array([0.5], np.float32), st)
    assert st["m"][0] == pytest.approx(0.0025, rel=1e-6)
    assert x[0] == pytest.approx(0.9, rel=1e-6)
    OO.rmsprop_step(x, np.array([0.5], np.float32), st)
    assert st["m"][0] == pytest.approx(0.004975, rel=1e-6)
    assert x[0] == pytest.approx(0.9 - 0.005 / (np.sqrt(0.004975) + 1e-8), rel=1e-6)
    # zero gradient: nothing moves, state decays
    y = np.array([2.0], np.float32)
    s2 = {"m": np.array([1.0], np.float32)}
    OO.rmsprop_step(y, np.zeros(1, np.float32), s2)
    assert y[0] == 2.0 and s2["m"][0] == np.float32(0.99)


def test_oracle_weight_decay_and_div():
    x = np.array([1.0, -2.0], np.float32)
    g = np.array([0.0, 0.0], np.float32)
    OO.rmsprop_step(x, g, {}, weightDecay=0.5, learningRate=0.1)
    # dfdx = wd * x; first step moves every weight by lr * sign(dfdx) / sqrt(1 - alpha) up to epsilon
    assert x[0] == pytest.approx(0.0, abs=1e-5)      # 1 - 0.1 * 0.5 / sqrt(0.01 * 0.25)
    assert x[1] == pytest.approx(-1.0, abs=1e-5)     # -2 + 0.1 * 1.0 / sqrt(0.01 * 1.0)
    assert np.array_equal(OO.gradient_div(np.array([3.0, 1.0], np.float32), 3), np.array([1.0, np.float32(1.0) / np.float32(3.0)], np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 3, 4, 1023, 1 << 20, (1 << 20) + 5])
@pytest.mark.parametrize("wd,div", [(0.0, 1.0), (0.0005, 256.0), (0.0, 37.0)])
def test_gpu_rmsprop_bit_exact(F, small_model, n, wd, div):
    """frcnn_rmsprop_step against the oracle on seeded buffers, three consecutive steps: integer-exact bar (every fp32
    result bit-identical: the kernel issues the same individually rounded operations, no FMA contraction)."""
    import torch
    rng = np.random.default_rng(n + int(div))
    w = rng.standard_normal(n).astype(np.float32)
    m = np.zeros(n, np.float32)
    wd_t, gd_t, md_t = torch.from_numpy(w.copy()).cuda(), None, torch.from_numpy(m.copy()).cuda()
    st = {}
    ffi, L = F.ffi, F.lib()
    for step in range(3):
        g = (rng.standard_normal(n) * 10.0 ** rng.integers(-6, 3, n)).astype(np.float32)
        if step == 1 and n > 2:
            g[:2] = 0.0
        gd_t = torch.from_numpy(g.copy()).cuda()
        rc = L.frcnn_rmsprop_step(small_model.ctx, ffi.cast("float*", wd_t.data_ptr()), ffi.cast("float*", gd_t.data_ptr()),
                                  ffi.cast("float*", md_t.data_ptr()), n, div, 1e-3, 0.99, 1e-8, wd)
        assert rc == 0, ffi.string(L.frcnn_last_error(small_model.ctx))
        gref = OO.gradient_div(g, div) if div != 1.0 else g
        OO.rmsprop_step(w, gref, st, learningRate=1e-3, alpha=0.99, epsilon=1e-8, weightDecay=wd)
        torch.cuda.synchronize()
        assert np.array_equal(wd_t.cpu().numpy().view(np.uint32), w.view(np.uint32))
        assert np.array_equal(md_t.cpu().numpy().view(np.uint32), st["m"].view(np.uint32))


@pytest.mark.gpu
def test_gpu_rmsprop_on_model(F, small_model):
    """The host mirror on the model's flat buffers: one step from a synthetic gradient moves `weights` exactly as the
    oracle does and refreshes the packed weights (pnet:forward changes)."""
    import torch
    from oracle import model as OM
    w0 = small_model.weights.clone()
    try:
        g = torch.randn_like(small_model.weights) * 1e-2
        small_model.gradient.copy_(g)
        st = {}
        F.rmsprop_step(small_model, st, learningRate=1e-4, grad_div=8.0)
        ref = w0.cpu().numpy().copy()
        OO.rmsprop_step(ref, OO.gradient_div(g.cpu().numpy(), 8.0), {}, learningRate=1e-4)
        assert np.array_equal(small_model.weights.cpu().numpy().view(np.uint32), ref.view(np.uint32))
        img = OM.synthetic_frame(122, 192, seed=1).cuda()
        a = small_model.pnet.forward(img)[4].clone()
        small_model.weights.copy_(w0)
        small_model.pack_weights()
        b = small_model.pnet.forward(img)[4]
        assert not torch.equal(a, b)
    finally:
        small_model.weights.copy_(w0)
        small_model.pack_weights()
