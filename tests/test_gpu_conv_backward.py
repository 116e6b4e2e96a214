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


def _first_wgrad(F, small_model, x, dy, pad, accumulate_into=None):
    n, _, h, w = x.shape
    dyd = dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    xd = x.contiguous().cuda()
    dw = torch.zeros(64, 3, 3, 3, device="cuda") if accumulate_into is None else accumulate_into
    ffi, L = F.ffi, F.lib()
    rc = L.frcnn_conv_first_wgrad(small_model.ctx, ffi.cast("const uint16_t*", dyd.data_ptr()),
                                  ffi.cast("const float*", xd.data_ptr()), n, h, w, pad, ffi.cast("float*", dw.data_ptr()), 1, ffi.NULL)
    assert rc == 0, ffi.string(L.frcnn_last_error(small_model.ctx))
    return dw


@pytest.mark.parametrize("n,h,w", [(1, 8, 32), (1, 16, 16), (2, 123, 77), (1, 450, 800), (3, 61, 96), (1, 5, 3)])
def test_conv_first_wgrad(F, small_model, n, h, w):
    """Weight gradient of the 3-channel first layer against autograd in float64.  The frame is NOT rounded: the kernel
    feeds it to the tensor cores as a hi + lo bf16 pair, so the only operand rounding is dy's (bf16 by contract) and
    the bar is fp32-accumulation tight: 1e-4 of the tensor maximum."""
    g = torch.Generator().manual_seed(n * 1000 + h + w)
    x = torch.randn(n, 3, h, w, generator=g, dtype=torch.float64).float().double().requires_grad_(False)
    wt = torch.zeros(64, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    dy = OM.bf16_round(torch.randn(n, 64, h, w, generator=g)).double()
    TF.conv2d(x, wt, None, padding=1).backward(dy)
    got = _first_wgrad(F, small_model, x.float(), dy.float(), 1).double().cpu()
    err = (got - wt.grad).abs().max().item()
    assert err <= 1e-4 * wt.grad.abs().max().item() + 1e-6, "max err %g at ref max %g" % (err, wt.grad.abs().max())


def test_conv_first_wgrad_exact_integers_and_accumulation(F, small_model):
    """Small-integer operands are exact in bf16 and in fp32 sums, so every tap / channel / swizzle address must come out
    bit-identical; a second call accumulates into the same tensor (accGradParameters semantics)."""
    g = torch.Generator().manual_seed(3)
    n, h, w = 2, 37, 45
    x = torch.randint(-3, 4, (n, 3, h, w), generator=g).float()
    dy = torch.randint(-2, 3, (n, 64, h, w), generator=g).float()
    wt = torch.zeros(64, 3, 3, 3, requires_grad=True)
    TF.conv2d(x, wt, None, padding=1).backward(dy)
    dw = _first_wgrad(F, small_model, x, dy, 1)
    assert torch.equal(dw.cpu(), wt.grad)
    dw = _first_wgrad(F, small_model, x, dy, 1, accumulate_into=dw)
    assert torch.equal(dw.cpu(), 2 * wt.grad)


def test_conv_first_wgrad_keeps_frame_precision(F, small_model):
    """A frame whose information sits below bf16 resolution (1 + tiny) still produces the fp32 answer: the lo half of
    the split carries it."""
    g = torch.Generator().manual_seed(9)
    n, h, w = 1, 40, 64
    x = (1.0 + torch.randn(n, 3, h, w, generator=g, dtype=torch.float64) * 2.0 ** -12).float()
    dy = OM.bf16_round(torch.randn(n, 64, h, w, generator=g))
    dy = dy - dy.mean(dim=(2, 3), keepdim=True)      # the constant part cancels: what is left comes from the tiny part
    dy = OM.bf16_round(dy)
    wt = torch.zeros(64, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    TF.conv2d(x.double(), wt, None, padding=1).backward(dy.double())
    got = _first_wgrad(F, small_model, x, dy, 1).double().cpu()
    ref = wt.grad
    hi_only = torch.zeros(64, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    TF.conv2d(OM.bf16_round(x).double(), hi_only, None, padding=1).backward(dy.double())
    err = (got - ref).abs().max().item()
    err_hi = (hi_only.grad - ref).abs().max().item()
    assert err <= 0.02 * err_hi + 1e-5, "split error %g, bf16-only frame would give %g" % (err, err_hi)
