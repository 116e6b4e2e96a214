"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/frcnn_b200.h declares,
its host-side geometry (Localizer / Anchors, exact double math) equals the oracle, and every compute entry point
fails loudly without a GPU (there is no CPU fallback).  No kernel is launched here."""
import ctypes
import os

import numpy as np
import pytest

from oracle import anchors as OA, localizer as OL, model as OM
from oracle.rect import Rect as ORect

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_library_exports_every_declared_symbol(F):
    names = F.declared_functions()
    assert len(names) >= 30 and "frcnn_detect" in names and "frcnn_nms" in names
    path = os.path.join(os.path.dirname(F.__file__), "libfrcnn_b200.so")
    so = ctypes.CDLL(path)
    for n in names:
        assert hasattr(so, n), "library does not export " + n
    assert F.lib().frcnn_version() == 100


def test_header_is_cdef_clean(F):
    # the header, minus preprocessor lines, is what LuaJIT ffi.cdef / cffi consume verbatim
    from frcnn_b200._lib import header_cdef
    src = header_cdef()
    assert "#" not in src and "extern" not in src
    for t in ("frcnn_block_desc", "frcnn_head_desc", "frcnn_fc_desc", "frcnn_candidate", "frcnn_detection"):
        assert t in src
    assert F.ffi.sizeof("frcnn_candidate") == 72 and F.ffi.sizeof("frcnn_detection") == 96


def test_no_cpu_fallback(F):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = F.ffi.new("frcnn_ctx**")
    assert F.lib().frcnn_create(ctx, 0, F.ffi.NULL) == 2  # FRCNN_E_CUDA
    assert b"no CPU fallback" in F.ffi.string(F.lib().frcnn_last_error(F.ffi.NULL))
    with pytest.raises(RuntimeError):
        F.vgg_small(F.duplo_cfg)
    m = F.vgg_small(F.duplo_cfg, device=-1)  # host-only context: geometry works, compute does not
    b = np.zeros((4, 4), np.float32)
    with pytest.raises(F.FrcnnError) as e:
        F.nms(b, 0.5, model=m)
    assert e.value.code == 2
    assert F.lib().frcnn_pack_weights(m.ctx) == 2


@pytest.mark.parametrize("which", ["small", "large"])
def test_host_geometry_equals_oracle(F, which):
    desc, cfg, fac, fcfg = (OM.VGG_SMALL, OM.CFG_DUPLO, F.vgg_small, F.duplo_cfg) if which == "small" else \
                           (OM.VGG_LARGE, OM.CFG_IMAGENET, F.vgg_large, F.imgnet_cfg)
    m = fac(fcfg, device=-1)
    # parameter table = oracle's param_specs (bind order is the flat-buffer order)
    specs = OM.param_specs(desc, cfg)
    assert m.param_names == [n for n, _ in specs]
    assert m.param_numel == [int(np.prod(s)) for _, s in specs]
    a = F.Anchors(m)
    oa = OA.Anchors(desc["layers"], desc["anchor_nets"], cfg["scales"])
    assert np.array_equal(a.w, oa.w) and np.array_equal(a.h, oa.h)  # bit-exact fp32 LUTs
    for i in range(4):
        assert a.localizers[i].layers == oa.localizers[i].layers
    loc = F.Localizer(m, 5)
    oloc = OL.Localizer(OL.trunk_layer_info(desc["layers"], 4))
    assert loc.layers == oloc.layers
    g = np.load(os.path.join(GOLD, "geometry_nms.npz"))
    want = g["roi_out_small"] if which == "small" else g["roi_out_large"]
    got = np.array([loc.inputToFeatureRect(F.Rect(*r)).unpack() for r in g["roi_in"]])
    assert np.array_equal(got, want)
    rng = np.random.default_rng(1)
    for _ in range(200):
        x0, y0 = rng.uniform(-100, 900), rng.uniform(-100, 500)
        r = (x0, y0, x0 + rng.uniform(0.5, 500), y0 + rng.uniform(0.5, 400))
        assert loc.inputToFeatureRect(F.Rect(*r)).unpack() == oloc.inputToFeatureRect(ORect(*r)).unpack()
        q = tuple(float(v) for v in rng.integers(0, 60, 4))
        assert loc.featureToInputRect(*q).unpack() == oloc.featureToInputRect(*q).unpack()
        # the optional layer_index argument (Localizer.lua:41-42,69-70): only the first k layers take part
        k = int(rng.integers(1, len(loc.layers) + 1))
        assert loc.inputToFeatureRect(F.Rect(*r), k).unpack() == oloc.inputToFeatureRect(ORect(*r), k).unpack()
        assert loc.featureToInputRect(*q, layer_index=k).unpack() == oloc.featureToInputRect(*q, layer_index=k).unpack()
    r = a.get(2, 3, 4, 5)
    o = oa.get(2, 3, 4, 5)
    assert r.unpack() == o.unpack() and r.index == o.index
    # head geometry is identical for the two models; the feature map carries the last block's filters
    assert m.output_dims(450, 800)[:2] == [(18, 55, 98), (18, 27, 48)]
    assert m.output_dims(450, 800)[4] == ((384, 29, 50) if which == "small" else (512, 29, 50))
    assert m.output_dims(600, 1000)[:4] == [(18, 73, 123), (18, 36, 61), (18, 34, 59), (18, 32, 57)]
    m.close()


def test_pnet_output_dims(F):
    m = F.vgg_small(F.duplo_cfg, device=-1)
    assert m.output_dims(450, 800) == [(18, 55, 98), (18, 27, 48), (18, 25, 46), (18, 23, 44), (384, 29, 50)]
    assert m.output_dims(122, 192) == [(18, 14, 22), (18, 6, 10), (18, 4, 8), (18, 2, 6), (384, 8, 12)]
    m.close()
    m = F.vgg_large(F.imgnet_cfg, device=-1)
    assert m.output_dims(600, 1000) == [(18, 73, 123), (18, 36, 61), (18, 34, 59), (18, 32, 57), (512, 38, 63)]
    m.close()


def test_plan_validation(F):
    m = F.vgg_small(F.duplo_cfg, device=-1)
    L, ffi = F.lib(), F.ffi
    # a second plan on the same ctx is a state error; bad localizer index is invalid
    scales = ffi.new("double[]", [32., 64., 128., 256.])
    blocks = ffi.new("frcnn_block_desc[]", 1)
    heads = ffi.new("frcnn_head_desc[]", 4)
    fcs = ffi.new("frcnn_fc_desc[]", 1)
    assert L.frcnn_model_plan(m.ctx, blocks, 1, heads, 4, fcs, 1, 16, 6, 6, scales, 4, -1.0) == 4
    n = ffi.new("int*")
    assert L.frcnn_localizer_layers(m.ctx, 9, ffi.NULL, 0, n) == 1
    m.close()


def test_nms_order_dispatch(F):
    # nms.lua:37-43: number -> column, 'area' -> area, anything else (incl. a score tensor) -> y2
    from frcnn_b200.nms import _order, ORDER_AREA, ORDER_COLUMN, ORDER_Y2
    assert _order(5) == (ORDER_COLUMN, 4)
    assert _order("area") == (ORDER_AREA, 0)
    assert _order(np.zeros(3)) == (ORDER_Y2, 0) and _order(None) == (ORDER_Y2, 0) and _order("score") == (ORDER_Y2, 0)


def test_example_record_layout_matches_header():
    """Model.pack_examples marshals frcnn_example[] as a numpy record array: field offsets and the record size must equal the
    C struct's (include/frcnn_b200.h), or frcnn_train_image / frcnn_train_batch would read garbage."""
    import importlib
    F = importlib.import_module("frcnn_b200")
    from frcnn_b200 import ffi
    dt = F.Model._EXAMPLE_DTYPE
    assert dt.itemsize == ffi.sizeof("frcnn_example")
    for name in ("anchor", "roi", "reg_target", "layer", "aspect", "y", "x", "class_index"):
        assert dt.fields[name][1] == ffi.offsetof("frcnn_example", name), name
    assert ffi.sizeof("frcnn_anchor_ref") == 16


def test_lua_glue_binds_only_declared_symbols_with_the_declared_arity():
    """The LuaJIT glue cannot run here (no Lua in the image), so it is checked statically against the header it cdef's:
    every C.frcnn_* call names a declared function and passes as many arguments as the declaration has parameters; every
    enum constant it reads exists."""
    import glob
    import os
    import re
    from frcnn_b200._lib import header_cdef
    cdef = header_cdef()
    decl = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*)\s+(frcnn_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", cdef, flags=re.S):
        args = m.group(2).strip()
        decl[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",")])
    consts = set(re.findall(r"\b(FRCNN_[A-Z0-9_]+)\b", cdef))
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "faster-rcnn.torch_b200", "lua")
    files = sorted(glob.glob(os.path.join(root, "*.lua")))
    assert len(files) >= 3
    calls = 0
    for path in files:
        src = open(path).read()
        src = re.sub(r"--[^\n]*", "", src)  # comments
        for m in re.finditer(r"\bC\.(FRCNN_[A-Z0-9_]+)", src):
            assert m.group(1) in consts, (path, m.group(1))
        for m in re.finditer(r"\bC\.(frcnn_[a-z0-9_]+)\s*(\()?", src):
            name = m.group(1)
            assert name in decl, "%s uses undeclared %s" % (os.path.basename(path), name)
            if not m.group(2):
                continue  # passed as a value (ffi.gc finaliser)
            depth, i, n_args, any_arg = 1, m.end(), 1, False
            while depth > 0:
                ch = src[i]
                if ch in "([{":
                    depth += 1
                elif ch in ")]}":
                    depth -= 1
                elif ch == "," and depth == 1:
                    n_args += 1
                if depth > 0 and not ch.isspace():
                    any_arg = True
                i += 1
            n = n_args if any_arg else 0
            assert n == decl[name], "%s: %s called with %d arguments, declared with %d" % (os.path.basename(path), name, n, decl[name])
            calls += 1
    assert calls >= 25


def test_lua_glue_installs_the_module_slot_overrides():
    """accelerate(model) must leave objective.lua / Detector.lua runnable unmodified: the glue derives the plan
    from the nn modules, binds weights and gradients, and installs pnet / cnet forward + backward and the `amp` class.
    Static check of the Lua source (no Lua in this image)."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "faster-rcnn.torch_b200", "lua")
    src = open(os.path.join(root, "frcnn_b200.lua")).read()
    code = re.sub(r"--[^\n]*", "", src)
    for slot in ("pnet.forward", "pnet.backward", "cnet.forward", "cnet.backward", "pnet.updateOutput", "cnet.updateOutput"):
        assert re.search(r"\b%s\s*=" % re.escape(slot), code), slot
    acc = code[code.index("function M.accelerate"):]
    acc = acc[:acc.index("\nend")]
    for call in ("derive_anchor_nets", "derive_class_layers", "frcnn_model_plan", "frcnn_bind_params", "M.pack(model)", "M.bind_grads(model)",
                 "M.install(model)", "M.install_amp(model)"):
        assert call in acc, call
    assert "nn.SpatialAdaptiveMaxPooling = cls" in code and "updateGradInput" in code and "self.indices" in code
    # the calls the reference's own files make on those slots, by the lines SURVEY 8b lists
    for fn in ("frcnn_pnet_forward_train", "frcnn_pnet_forward", "frcnn_pnet_backward", "frcnn_cnet_forward_train", "frcnn_cnet_forward",
               "frcnn_cnet_backward", "frcnn_adaptive_maxpool_forward", "frcnn_adaptive_maxpool_backward", "frcnn_train_batch",
               "frcnn_dp_init_all", "frcnn_dp_allreduce", "frcnn_find_positive", "frcnn_sample_negative", "frcnn_find_nearby_negative",
               "frcnn_rmsprop_step", "frcnn_detect_begin", "frcnn_detect_end"):
        assert "C.%s(" % fn in code, fn
