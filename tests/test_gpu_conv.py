"""GPU tests of the tcgen05 implicit-GEMM convolution (csrc/conv_igemm.cu) and of pnet:forward built on it.

Floating point: the tensor cores take bf16 operands and accumulate in fp32.  Kernel-level tests compare with a
PyTorch fp32 convolution of the SAME bf16-rounded operands (difference = accumulation order + one bf16 rounding of
the output: rtol 1e-2 / atol scaled to the output magnitude).  pnet tests compare with the fp32 oracle, with the
tolerance stated per test."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

from oracle import model as OM

pytestmark = pytest.mark.gpu


def _conv_case(F, model, n, h, w, cin, cout, k, pad, splits=0, bn=0, prelu=True, scale=1.0, seed=0, mt=0, pool=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    slope = torch.tensor([0.2])
    xq, wq = OM.bf16_round(x), OM.bf16_round(wt)
    ref = TF.conv2d(xq, wq, bias, padding=pad)
    if prelu:
        ref = torch.where(ref > 0, ref, ref * slope)
    ref = ref * scale
    x_nhwc = xq.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
    if pool:  # nn.SpatialMaxPooling(2,2,2,2):ceil() fused into the epilogue
        ref = TF.max_pool2d(ref, 2, 2, ceil_mode=True)
        ho, wo = (ho + 1) // 2, (wo + 1) // 2
    out = torch.full((n, ho, wo, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    wd, bd, sd = wt.cuda(), bias.cuda(), slope.cuda()
    ffi, L = F.ffi, F.lib()
    ms = ffi.new("float*")
    rc = L.frcnn_conv_bf16(model.ctx, ffi.cast("const uint16_t*", x_nhwc.data_ptr()), ffi.cast("const float*", wd.data_ptr()),
                           ffi.cast("const float*", bd.data_ptr()),
                           ffi.cast("const float*", sd.data_ptr()) if prelu else ffi.NULL, scale, n, h, w, cin, cout, k, pad,
                           splits, bn, mt, 1 if pool else 0, ffi.cast("uint16_t*", out.data_ptr()), 1, ms)
    assert rc == 0, ffi.string(L.frcnn_last_error(model.ctx))
    got = out.float().cpu().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    tol = 1e-2 * ref.abs() + 1e-2 * ref.abs().max()
    assert bool((err <= tol).all()), "max err %g at ref max %g" % (err.max(), ref.abs().max())
    return got, ref


@pytest.mark.parametrize("case", [
    # n, h, w, cin, cout, k, pad
    (1, 16, 16, 64, 64, 3, 1), (1, 20, 37, 64, 128, 3, 1), (2, 29, 50, 128, 128, 3, 1), (1, 57, 100, 128, 256, 3, 1),
    (1, 29, 50, 256, 384, 3, 1), (1, 29, 50, 384, 384, 3, 1), (1, 57, 100, 256, 256, 3, 0), (1, 29, 50, 384, 256, 5, 0),
    (1, 29, 50, 384, 256, 7, 0), (1, 1, 300, 512, 256, 1, 0), (3, 9, 7, 64, 192, 3, 1), (1, 113, 200, 64, 128, 3, 1),
])
def test_conv_shapes(F, small_model, case):
    _conv_case(F, small_model, *case, seed=sum(case))


@pytest.mark.parametrize("bn", [64, 128, 192, 256])
def test_conv_tile_widths(F, small_model, bn):
    cout = {64: 64, 128: 256, 192: 384, 256: 512}[bn]
    _conv_case(F, small_model, 1, 30, 41, 128, cout, 3, 1, bn=bn, seed=bn)


@pytest.mark.parametrize("splits", [2, 3, 9])
def test_conv_split_k(F, small_model, splits):
    _conv_case(F, small_model, 1, 27, 48, 384, 256, 3, 0, splits=splits, seed=splits)


def test_conv_epilogue_variants(F, small_model):
    _conv_case(F, small_model, 1, 24, 24, 64, 64, 3, 1, prelu=False)
    _conv_case(F, small_model, 1, 24, 24, 64, 64, 3, 1, scale=0.6)


@pytest.mark.parametrize("case", [
    # n, h, w, cin, cout, k, pad  -- even / odd map sizes (ceil-mode windows clipped at the border), all tile widths
    (1, 16, 16, 64, 64, 3, 1), (1, 29, 51, 128, 128, 3, 1), (2, 57, 100, 128, 256, 3, 1), (1, 113, 200, 64, 128, 3, 1),
    (1, 57, 99, 256, 384, 3, 1), (1, 225, 400, 64, 128, 3, 1), (3, 9, 7, 64, 192, 3, 1), (1, 75, 125, 256, 512, 3, 1),
])
def test_conv_fused_pool(F, small_model, case):
    _conv_case(F, small_model, *case, seed=sum(case), pool=True)


@pytest.mark.parametrize("pool", [False, True])
@pytest.mark.parametrize("case", [(1, 64, 64, 64, 128, 3, 1), (2, 113, 200, 128, 128, 3, 1), (1, 45, 77, 64, 64, 3, 1)])
def test_conv_two_subtiles_per_cta(F, small_model, case, pool):
    """MT = 2: CTA tiles of 256 pixels sharing one weight tile (the narrow, L2-bound layers)."""
    _conv_case(F, small_model, *case, seed=sum(case) + 1, mt=2, pool=pool)


@pytest.mark.parametrize("pool", [False, True])
@pytest.mark.parametrize("mt", [11, 12])
@pytest.mark.parametrize("case,bn", [
    ((1, 16, 16, 64, 64, 3, 1), 0), ((1, 29, 51, 128, 128, 3, 1), 128), ((2, 57, 100, 128, 256, 3, 1), 256),
    ((1, 57, 99, 256, 384, 3, 1), 192), ((1, 57, 99, 256, 384, 3, 1), 128), ((1, 45, 77, 64, 128, 3, 1), 64),
    ((1, 33, 20, 128, 256, 3, 0), 0), ((1, 40, 41, 64, 64, 2, 0), 0), ((4, 113, 200, 128, 256, 3, 1), 0),
])
def test_conv_halo_kernel(F, small_model, case, bn, mt, pool):
    """conv_halo_kernel (mt = 11 / 12 forces it with 1 / 2 sub-tiles): the CTA tile + halo is ONE TMA box per 64-channel
    chunk and every filter tap is a row-shifted UMMA descriptor into it.  Covers every tile width, the single-accumulator
    configuration (bn 256 x 2 sub-tiles), several units per persistent CTA (ring phase wrap), odd map sizes, pad 0."""
    _conv_case(F, small_model, *case, seed=sum(case) + mt, bn=bn, mt=mt, pool=pool)


@pytest.mark.parametrize("pool", [False, True])
@pytest.mark.parametrize("mt", [21, 22])
@pytest.mark.parametrize("case,bn", [
    ((1, 16, 16, 64, 64, 3, 1), 0), ((1, 29, 51, 128, 128, 3, 1), 128), ((2, 57, 100, 128, 256, 3, 1), 256),
    ((1, 57, 99, 256, 384, 3, 1), 192), ((1, 57, 99, 256, 384, 3, 1), 128), ((1, 45, 77, 64, 128, 3, 1), 64),
    ((1, 33, 20, 128, 256, 3, 0), 0), ((1, 40, 41, 64, 64, 2, 0), 0), ((4, 113, 200, 128, 256, 3, 1), 0),
    ((1, 113, 200, 128, 256, 3, 1), 0), ((1, 225, 400, 64, 128, 3, 1), 0), ((1, 57, 100, 256, 384, 3, 1), 0),
])
def test_conv_pair_kernel(F, small_model, case, bn, mt, pool):
    """conv_pair_kernel (mt = 21 / 22): the halo kernel on CTA pairs -- tcgen05.mma cta_group::2 with M = 256, each CTA of the
    pair loading its own pixel tile and half of every weight box.  Same coverage as the halo kernel plus the headline
    layer shapes: every tile width, ragged pair tiles (the second CTA's tile partly or wholly outside the map), several
    units per pair (ring / accumulator phase wrap), pad 0, 2x2 filters, the fused pool."""
    _conv_case(F, small_model, *case, seed=sum(case) + mt, bn=bn, mt=mt, pool=pool)


@pytest.mark.parametrize("pool", [False, True])
@pytest.mark.parametrize("case,bn,mt", [
    ((1, 29, 51, 128, 128, 3, 1), 128, 31), ((1, 57, 99, 256, 384, 3, 1), 192, 31), ((1, 45, 77, 64, 128, 3, 1), 64, 31),
    ((2, 57, 100, 128, 256, 3, 1), 256, 41), ((1, 57, 99, 256, 384, 3, 1), 192, 41), ((4, 113, 200, 128, 256, 3, 1), 0, 41),
    ((1, 225, 400, 64, 128, 3, 1), 0, 41), ((1, 45, 77, 64, 128, 3, 1), 64, 42), ((1, 113, 200, 128, 256, 3, 1), 128, 31),
])
def test_conv_two_ctas_per_sm(F, small_model, case, bn, mt, pool):
    """The halo kernel (mt = 31) and the CTA-pair kernel (mt = 41 / 42) sized for TWO resident CTAs per SM: half the shared
    memory and TMEM columns per CTA, two-slot activation ring, <= 85 registers -- same arithmetic, shallower pipeline."""
    _conv_case(F, small_model, *case, seed=sum(case) + mt, bn=bn, mt=mt, pool=pool)


@pytest.mark.parametrize("pool", [False, True])
@pytest.mark.parametrize("case", [
    (1, 16, 16, 64, 64, 3, 1), (1, 29, 51, 128, 128, 3, 1), (1, 45, 77, 64, 128, 3, 1), (1, 225, 400, 64, 128, 3, 1),
    (1, 225, 400, 128, 128, 3, 1), (8, 113, 200, 128, 128, 3, 1), (1, 33, 20, 128, 64, 3, 0), (3, 61, 96, 64, 128, 3, 1),
])
def test_conv_pair_resident_weights(F, small_model, case, pool):
    """conv_pair_bres_kernel (mt = 51): CTA pairs with every weight box of the layer resident in shared memory (loaded while
    the pair's first unit runs, reused by all later units) -- the Cout = 64 / 128 layers (conv2_x at their real sizes, many
    units per pair, ragged pair tiles, pad 0, the fused pool); bit-identical to the streaming pair kernel."""
    a, _ = _conv_case(F, small_model, *case, seed=sum(case) + 51, bn=0, mt=51, pool=pool)
    b, _ = _conv_case(F, small_model, *case, seed=sum(case) + 51, bn=case[4], mt=21, pool=pool)
    assert torch.equal(a, b)


@pytest.mark.parametrize("pool", [False, True])
@pytest.mark.parametrize("case", [
    (1, 16, 16, 64, 128, 3, 1), (1, 29, 51, 128, 128, 3, 1), (1, 45, 77, 64, 128, 3, 1), (1, 225, 400, 64, 128, 3, 1),
    (1, 225, 400, 128, 128, 3, 1), (2, 57, 100, 256, 384, 3, 1), (1, 57, 99, 384, 384, 3, 1), (1, 33, 20, 128, 256, 3, 0),
    (3, 61, 96, 64, 128, 3, 1), (1, 113, 200, 128, 256, 3, 1),
])
def test_conv_swapped_operands(F, small_model, case, pool):
    """conv_halo_kernel<128, 2, 3, 1, SWAP> (mt = 61): the 128-filter weight box as the MMA's A operand, an 8 x 32-pixel halo
    tile as its 256-wide B operand, accumulator lanes = output channels, pooling in registers, transposed staging.  Same
    products in the same K order as the halo kernel: bit-identical outputs (ragged tiles, several filter tiles, pad 0,
    the fused pool, the headline conv2_x / conv4_x shapes)."""
    a, _ = _conv_case(F, small_model, *case, seed=sum(case) + 61, bn=0, mt=61, pool=pool)
    b, _ = _conv_case(F, small_model, *case, seed=sum(case) + 61, bn=128, mt=12, pool=pool)
    assert torch.equal(a, b)


def test_conv_pair_equals_halo_kernel(F, small_model):
    """Same products, same fp32 accumulation order per output (chunk-major, taps inside): the pair kernel's outputs are
    bit-identical to the single-CTA halo kernel's."""
    a, _ = _conv_case(F, small_model, 1, 57, 100, 128, 256, 3, 1, seed=5, mt=12)
    b, _ = _conv_case(F, small_model, 1, 57, 100, 128, 256, 3, 1, seed=5, mt=22)
    assert torch.equal(a, b)


def test_conv_halo_equals_tap_kernel(F, small_model):
    """Both kernels accumulate the same products in fp32 in the same (chunk-major vs tap-major) grouping only up to
    fp32 rounding: outputs agree to one bf16 ulp, and exactly on small-integer data."""
    a, _ = _conv_case(F, small_model, 1, 57, 100, 128, 256, 3, 1, seed=5, mt=1)
    b, _ = _conv_case(F, small_model, 1, 57, 100, 128, 256, 3, 1, seed=5, mt=12)
    assert (a - b).abs().max().item() <= 2.0 ** -7 * a.abs().max().item()


@pytest.mark.parametrize("pool", [False, True])
@pytest.mark.parametrize("n,h,w", [(1, 16, 16), (1, 450, 800), (2, 123, 77), (1, 61, 96)])
def test_conv_first_layer(F, small_model, n, h, w, pool):
    """The fused first layer (in-kernel im2col of the fp32 NCHW frame, K = 27) against conv2d on bf16-rounded
    operands."""
    g = torch.Generator().manual_seed(h + w + n)
    x = torch.randn(n, 3, h, w, generator=g)
    wt = torch.randn(64, 3, 3, 3, generator=g) * (2.0 / 27) ** 0.5
    bias = torch.randn(64, generator=g) * 0.1
    slope = torch.tensor([0.2])
    ref = TF.conv2d(OM.bf16_round(x), OM.bf16_round(wt), bias, padding=1)
    ref = torch.where(ref > 0, ref, ref * slope)
    ho, wo = h, w
    if pool:
        ref = TF.max_pool2d(ref, 2, 2, ceil_mode=True)
        ho, wo = (h + 1) // 2, (w + 1) // 2
    out = torch.full((n, ho, wo, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    xd, wd, bd, sd = x.cuda(), wt.cuda(), bias.cuda(), slope.cuda()
    ffi, L = F.ffi, F.lib()
    rc = L.frcnn_conv_first(small_model.ctx, ffi.cast("const float*", xd.data_ptr()), ffi.cast("const float*", wd.data_ptr()),
                            ffi.cast("const float*", bd.data_ptr()), ffi.cast("const float*", sd.data_ptr()), 1.0, n, h, w, 64, 1,
                            1 if pool else 0, ffi.cast("uint16_t*", out.data_ptr()), 1, ffi.NULL)
    assert rc == 0, ffi.string(L.frcnn_last_error(small_model.ctx))
    got = out.float().cpu().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    tol = 1e-2 * ref.abs() + 1e-2 * ref.abs().max()
    assert bool((err <= tol).all()), "max err %g at ref max %g" % (err.max(), ref.abs().max())


@pytest.mark.parametrize("mt", [1, 11, 12])
def test_conv_exact_integers(F, small_model, mt):
    """Small-integer operands are exact in bf16 and in fp32 accumulation: the result must be bit-identical to the
    reference convolution -- catches any tap / channel / swizzle mis-addressing that tolerances could hide."""
    g = torch.Generator().manual_seed(1)
    n, h, w, cin, cout, k, pad = 1, 23, 35, 128, 128, 3, 1
    x = torch.randint(-2, 3, (n, cin, h, w), generator=g).float()
    wt = torch.randint(-2, 3, (cout, cin, k, k), generator=g).float()
    bias = torch.randint(-3, 4, (cout,), generator=g).float()
    ref = TF.conv2d(x, wt, bias, padding=pad)
    ref = ref.clamp(-256, 256)  # keep outputs exactly representable in bf16
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    out = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    wd, bd = wt.cuda(), bias.cuda()
    ffi, L = F.ffi, F.lib()
    rc = L.frcnn_conv_bf16(small_model.ctx, ffi.cast("const uint16_t*", x_nhwc.data_ptr()), ffi.cast("const float*", wd.data_ptr()),
                           ffi.cast("const float*", bd.data_ptr()), ffi.NULL, 1.0, n, h, w, cin, cout, k, pad, 0, 0, mt, 0,
                           ffi.cast("uint16_t*", out.data_ptr()), 1, ffi.NULL)
    assert rc == 0
    got = out.float().cpu().permute(0, 3, 1, 2)
    mask = (TF.conv2d(x, wt, bias, padding=pad).abs() <= 256)
    assert torch.equal(got[mask], ref[mask])


# per-element tolerance of pnet:forward against the PURE fp32 oracle: |cuda - fp32| <= atol + rtol * |fp32|
PNET_TOL = {"fp16": dict(atol=6e-3, rtol=4e-3), "bf16": dict(atol=4e-2, rtol=2e-2)}


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("h,w", [(122, 192), (450, 800)])
def test_pnet_forward_vs_oracle(F, small_model, h, w, precision):
    """pnet:forward (model_utilities.lua:3-58), evaluate mode, against the oracle at the headline size too.  Stated
    tolerance per element against the pure fp32 oracle (the reference's arithmetic): PNET_TOL; against the oracle run
    with the same 16-bit operand rounding, half of that (the stored intermediate activations add one rounding per
    layer).  fp16 operands (the default) are 8x closer to the reference than bf16."""
    img = OM.synthetic_frame(h, w, seed=1)
    small_model.set_eval_precision(precision)
    try:
        outs = small_model.pnet.forward(img.cuda())
    finally:
        small_model.set_eval_precision("fp16")
    q = OM.fp16_round if precision == "fp16" else OM.bf16_round
    with torch.no_grad():
        want_q = OM.pnet_forward(OM.VGG_SMALL, small_model.oracle_params, img, quant=q)
        want_f = OM.pnet_forward(OM.VGG_SMALL, small_model.oracle_params, img)
    tol = PNET_TOL[precision]
    for i, (o, wq, wf) in enumerate(zip(outs, want_q, want_f)):
        assert tuple(o.shape) == tuple(wq.shape)
        o = o.cpu()
        assert bool(((o - wf).abs() <= tol["atol"] + tol["rtol"] * wf.abs()).all()), ("fp32 oracle, output %d" % i, float((o - wf).abs().max()))
        assert bool(((o - wq).abs() <= 0.5 * (tol["atol"] + tol["rtol"] * wq.abs())).all()), ("rounded oracle, output %d" % i, float((o - wq).abs().max()))


def test_pnet_batch_equals_single(F, small_model):
    """Images of a batch are independent (objective.lua:65): batched forward == per-image forward -- bit for bit on
    the trunk; the anchor-head maps are summed from split-K slices whose split factor depends on the batch size,
    so they agree to fp32 rounding (1e-5 of the map magnitude)."""
    imgs = torch.stack([OM.synthetic_frame(122, 192, seed=s) for s in range(3)]).cuda()
    outs = small_model.pnet.forward(imgs)
    for s in range(3):
        single = small_model.pnet.forward(imgs[s])
        assert torch.equal(outs[4][s], single[4])
        for a, b in zip(outs[:4], single[:4]):
            assert (a[s] - b).abs().max().item() <= 1e-5 * b.abs().max().item()


@pytest.mark.parametrize("n,h,w", [(1, 122, 192), (1, 450, 800), (3, 122, 192)])
def test_pnet_forward_throughput_schedule(F, small_model, n, h, w):
    """FRCNN_SCHED_THROUGHPUT: the four AnchorNetworks as unsplit halo-kernel units with bias + PReLU + the 1x1 conv fused
    into the epilogue (fp32, fixed order).  Same sums as the split-K + tail-kernel schedule in another order: the 18-channel
    maps agree to fp32 rounding (1e-4 of the map magnitude: up to 18 816 products per sum), the trunk is untouched (bit-identical), and the fused path
    is itself bit-reproducible and batch-invariant (no split factor depends on the batch)."""
    imgs = torch.stack([OM.synthetic_frame(h, w, seed=7 + s) for s in range(n)]).cuda()
    base = [o.clone() for o in small_model.pnet.forward(imgs)]
    small_model.set_schedule("throughput")
    try:
        fused = [o.clone() for o in small_model.pnet.forward(imgs)]
        again = small_model.pnet.forward(imgs)
        for a, b in zip(fused, again):
            assert torch.equal(a, b)
        assert torch.equal(fused[4], base[4])
        for a, b in zip(fused[:4], base[:4]):
            assert a.shape == b.shape
            assert (a - b).abs().max().item() <= 1e-4 * b.abs().max().item()
        if n > 1:
            for s in range(n):
                single = small_model.pnet.forward(imgs[s])
                for a, b in zip(fused[:4], single[:4]):
                    assert torch.equal(a[s], b)
    finally:
        small_model.set_schedule("latency")


def test_vgg_large_throughput_schedule(F):
    m = F.vgg_large(F.imgnet_cfg)
    p = OM.init_params(OM.VGG_LARGE, OM.CFG_IMAGENET, seed=2, randomize_aux=True)
    m.load_params(p)
    m.set_schedule("throughput")
    img = OM.synthetic_frame(150, 200, seed=4)
    outs = m.pnet.forward(img.cuda())
    with torch.no_grad():
        want = OM.pnet_forward(OM.VGG_LARGE, p, img)   # pure fp32
    for o, q in zip(outs, want):
        assert tuple(o.shape) == tuple(q.shape)
        assert bool(((o.cpu() - q).abs() <= PNET_TOL["fp16"]["atol"] + PNET_TOL["fp16"]["rtol"] * q.abs()).all()), float((o.cpu() - q).abs().max())
    m.close()


def test_pnet_forward_is_deterministic(F, small_model):
    """Every kernel of pnet:forward has a fixed summation order (split-K slices, no atomics): two runs on the same
    frame are bit-identical, which is what makes the decode / NMS index parity reproducible end to end."""
    img = OM.synthetic_frame(450, 800, seed=5).cuda()
    a = [o.clone() for o in small_model.pnet.forward(img)]
    b = small_model.pnet.forward(img)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_vgg_large_forward(F):
    m = F.vgg_large(F.imgnet_cfg)
    p = OM.init_params(OM.VGG_LARGE, OM.CFG_IMAGENET, seed=2, randomize_aux=True)
    m.load_params(p)
    img = OM.synthetic_frame(150, 200, seed=4)
    outs = m.pnet.forward(img.cuda())
    with torch.no_grad():
        want = OM.pnet_forward(OM.VGG_LARGE, p, img)   # pure fp32
    for o, q in zip(outs, want):
        assert tuple(o.shape) == tuple(q.shape)
        assert bool(((o.cpu() - q).abs() <= PNET_TOL["fp16"]["atol"] + PNET_TOL["fp16"]["rtol"] * q.abs()).all()), float((o.cpu() - q).abs().max())
    m.close()
