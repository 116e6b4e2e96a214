"""Host mirrors of Localizer.lua and Anchors.lua.  The arithmetic lives in the C library (exact double math on the
host side of libfrcnn_b200.so); these classes keep the reference's method names."""
import math

import numpy as np

from ._lib import check, ffi, lib
from .rect import Rect


class Localizer:
    """Localizer.new(pnet.outnode.children[i]) (Localizer.lua:6-39).  `which` is the 1-based output index of pnet:
    1..#anchor_nets for the anchor heads, #anchor_nets+1 for the last conv block (Detector.lua:12)."""

    def __init__(self, model, which):
        self.model, self.which = model, which - 1
        n = ffi.new("int*")
        check(model.ctx, lib().frcnn_localizer_layers(model.ctx, self.which, ffi.NULL, 0, n))
        buf = ffi.new("int[]", 6 * n[0])
        check(model.ctx, lib().frcnn_localizer_layers(model.ctx, self.which, buf, n[0], n))
        keys = ("kW", "kH", "dW", "dH", "padW", "padH")
        self.layers = [dict(zip(keys, [buf[6 * i + j] for j in range(6)])) for i in range(n[0])]

    def inputToFeatureRect(self, rect, layer_index=None):
        if layer_index is not None and layer_index != len(self.layers):
            raise NotImplementedError("partial layer ranges are not used on the detection path")
        src = ffi.new("double[4]", list(rect.unpack()))
        dst = ffi.new("double[4]")
        check(self.model.ctx, lib().frcnn_input_to_feature_rect(self.model.ctx, self.which, src, dst))
        return Rect(dst[0], dst[1], dst[2], dst[3])

    def featureToInputRect(self, minX, minY, maxX, maxY, layer_index=None):
        if layer_index is not None and layer_index != len(self.layers):
            raise NotImplementedError("partial layer ranges are not used on the detection path")
        src = ffi.new("double[4]", [minX, minY, maxX, maxY])
        dst = ffi.new("double[4]")
        check(self.model.ctx, lib().frcnn_feature_to_input_rect(self.model.ctx, self.which, src, dst))
        return Rect(dst[0], dst[1], dst[2], dst[3])


class Anchors:
    """Anchors.new(pnet, scales) (Anchors.lua:7-58): fp32 LUTs w/h [scale][aspect][200][{min,max}]."""

    def __init__(self, model, scales=None):
        self.model = model
        scales = scales or model.cfg["scales"]
        self.localizers = [Localizer(model, i + 1) for i in range(len(scales))]
        n = len(scales) * 3 * 200 * 2
        w, h = ffi.new("float[]", n), ffi.new("float[]", n)
        check(model.ctx, lib().frcnn_anchors_build(model.ctx, w, h))
        self.w = np.frombuffer(ffi.buffer(w), dtype=np.float32).reshape(len(scales), 3, 200, 2).copy()
        self.h = np.frombuffer(ffi.buffer(h), dtype=np.float32).reshape(len(scales), 3, 200, 2).copy()

    def get(self, layer, aspect, y, x):  # Anchors.lua:60-67, 1-based indices
        w, h = self.w, self.h
        r = Rect(w[layer - 1, aspect - 1, x - 1, 0], h[layer - 1, aspect - 1, y - 1, 0],
                 w[layer - 1, aspect - 1, x - 1, 1], h[layer - 1, aspect - 1, y - 1, 1])
        r.layer, r.aspect = layer, aspect
        r.index = ((aspect * 6 - 5, aspect * 6), y, x)
        return r

    @staticmethod
    def inputToAnchor(anchor, rect):  # Anchors.lua:237-243
        return np.array([(rect.minX - anchor.minX) / anchor.width(), (rect.minY - anchor.minY) / anchor.height(),
                         math.log(rect.width() / anchor.width()), math.log(rect.height() / anchor.height())],
                        dtype=np.float32)

    @staticmethod
    def anchorToInput(anchor, t):  # Anchors.lua:245-252
        t = [float(v) for v in t]
        return Rect.fromXYWidthHeight(t[0] * anchor.width() + anchor.minX, t[1] * anchor.height() + anchor.minY,
                                      math.exp(t[2]) * anchor.width(), math.exp(t[3]) * anchor.height())
