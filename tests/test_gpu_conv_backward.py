"""GPU tests of the convolution backward primitives (pnet:backward, objective.lua:189) against torch.autograd on the
same bf16-rounded operands.  Tolerance: 1 % of the tensor maximum + 1 % relative (bf16 operands, fp32 accumulation;
dgrad output is rounded to bf16, wgrad stays fp32)."""
import pytest
import torch
import torch.nn.functional as TF

from oracle import model as OM

pytestmark = pytest.mark.gpu

CASES = [
    # n, h, w, cin, cout, k, pad
    (1, 16, 16, 64, 64, 3, 1), (2, 29, 50, 128, 256, 3, 1), (1, 57, 100, 256, 256, 3, 1), (1, 29, 50, 384, 256, 5, 0),
    (1, 29, 50, 384, 256, 7, 0), (1, 57, 100, 256, 256, 3, 0), (3, 9, 7, 64, 192, 3, 1), (1, 113, 200, 64, 128, 3, 1),
    (1, 30, 41, 128, 64, 1, 0), (1, 29, 50, 384, 384, 3, 1), (4, 57, 100, 128, 128, 3, 1), (1, 40, 40, 64, 64, 2, 0),
]


def _ref(case, seed):
    n, h, w, cin, cout, k, pad = case
    g = torch.Generator().manual_seed(seed)
    x = OM.bf16_round(torch.randn(n, cin, h, w, generator=g)).requires_grad_(True)
    wt = OM.bf16_round(torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5).requires_grad_(True)
    y = TF.conv2d(x, wt, None, padding=pad)
    dy = OM.bf16_round(torch.randn(y.shape, generator=g))
    y.backward(dy)
    return x.detach(), wt.detach(), dy, x.grad, wt.grad


def _close(got, ref, what):
    err = (got - ref).abs()
    tol = 1e-2 * ref.abs() + 1e-2 * ref.abs().max()
    assert bool((err <= tol).all()), "%s: max err %g at ref max %g" % (what, err.max(), ref.abs().max())


@pytest.mark.parametrize("case", CASES)
def test_conv_dgrad(F, small_model, case):
    n, h, w, cin, cout, k, pad = case
    x, wt, dy, dx_ref, _ = _ref(case, sum(case))
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    wd = wt.cuda()
    ffi, L = F.ffi, F.lib()
    rc = L.frcnn_conv_dgrad_bf16(small_model.ctx, ffi.cast("const uint16_t*", dy_nhwc.data_ptr()), ffi.cast("const float*", wd.data_ptr()),
                                 n, h, w, cin, cout, k, pad, ffi.cast("uint16_t*", dx.data_ptr()))
    assert rc == 0, ffi.string(L.frcnn_last_error(small_model.ctx))
    _close(dx.float().cpu().permute(0, 3, 1, 2), dx_ref, "dgrad")


@pytest.mark.parametrize("case", CASES)
def test_conv_wgrad(F, small_model, case):
    n, h, w, cin, cout, k, pad = case
    x, wt, dy, _, dw_ref = _ref(case, sum(case) + 1)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    base = torch.randn(cout, cin, k, k)  # wgrad ACCUMULATES into the gradient buffer (objective.lua:49,189)
    dw = base.clone().cuda()
    ffi, L = F.ffi, F.lib()
    rc = L.frcnn_conv_wgrad_bf16(small_model.ctx, ffi.cast("const uint16_t*", x_nhwc.data_ptr()),
                                 ffi.cast("const uint16_t*", dy_nhwc.data_ptr()), n, h, w, cin, cout, k, pad,
                                 ffi.cast("float*", dw.data_ptr()))
    assert rc == 0, ffi.string(L.frcnn_last_error(small_model.ctx))
    _close(dw.cpu() - base, dw_ref, "wgrad")
