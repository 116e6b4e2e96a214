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

    def inputToFeatureRect(self, rect, layer_index=None):  # Localizer.lua:41-67
        """layer_index (Localizer.lua:42): only the first layer_index layers; default: all of them."""
        src = ffi.new("double[4]", list(rect.unpack()))
        dst = ffi.new("double[4]")
        check(self.model.ctx, lib().frcnn_input_to_feature_rect_upto(self.model.ctx, self.which, int(layer_index or 0), src, dst))
        return Rect(dst[0], dst[1], dst[2], dst[3])

    def featureToInputRect(self, minX, minY, maxX, maxY, layer_index=None):  # Localizer.lua:69-79
        src = ffi.new("double[4]", [minX, minY, maxX, maxY])
        dst = ffi.new("double[4]")
        check(self.model.ctx, lib().frcnn_feature_to_input_rect_upto(self.model.ctx, self.which, int(layer_index or 0), src, dst))
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
        # centre-point bins for findNearby (Anchors.lua:22-30,46,54): key = floor(centre / BIN_SIZE) -> {scale, aspect, cell}
        self.cx, self.cy = {}, {}
        for i, loc in enumerate(self.localizers):
            for j in range(3):
                for v in range(1, 201):
                    cy = loc.featureToInputRect(0, v - 1, 0, v).center()[1]
                    cx = loc.featureToInputRect(v - 1, 0, v, 0).center()[0]
                    self.cy.setdefault(math.floor(cy / self.BIN_SIZE), []).append((i + 1, j + 1, v))
                    self.cx.setdefault(math.floor(cx / self.BIN_SIZE), []).append((i + 1, j + 1, v))

    BIN_SIZE = 16  # Anchors.lua:5

    def findNearby(self, centerX, centerY):  # Anchors.lua:69-84 (host-side, like the reference)
        found = []
        xl = self.cx.get(math.floor(centerX / self.BIN_SIZE))
        yl = self.cy.get(math.floor(centerY / self.BIN_SIZE))
        if xl and yl:
            for y in yl:
                for x in xl:
                    if y[0] == x[0] and y[1] == x[1]:
                        found.append(self.get(y[0], y[1], y[2], x[2]))
        return found

    def findRangesXY(self, rect, clip_rect=None):  # Anchors.lua:86-145 (host-side; the device twin lives in label_kernels.cu)
        """Per (scale, aspect) the 1-based LUT cell ranges [lx, ux) x [ly, uy) of the anchors that touch `rect` (and lie
        inside `clip_rect`): list of dicts {layer, aspect, lx, ly, ux, uy, xs, ys} like the Lua tables."""
        def lower_bound(t, value):   # first 1-based index with t[i] >= value
            return int(np.searchsorted(t, value, side="left")) + 1

        def upper_bound(t, value):   # first 1-based index with t[i] > value
            return int(np.searchsorted(t, value, side="right")) + 1

        ranges = []
        w, h = self.w.astype(np.float64), self.h.astype(np.float64)
        for i in range(w.shape[0]):
            for j in range(3):
                lx, ly = upper_bound(w[i, j, :, 1], rect.minX), upper_bound(h[i, j, :, 1], rect.minY)
                ux, uy = lower_bound(w[i, j, :, 0], rect.maxX), lower_bound(h[i, j, :, 0], rect.maxY)
                if clip_rect is not None:
                    lx, ly = max(lx, lower_bound(w[i, j, :, 0], clip_rect.minX)), max(ly, lower_bound(h[i, j, :, 0], clip_rect.minY))
                    ux, uy = min(ux, upper_bound(w[i, j, :, 1], clip_rect.maxX)), min(uy, upper_bound(h[i, j, :, 1], clip_rect.maxY))
                if ux > lx and uy > ly:
                    ranges.append(dict(layer=i + 1, aspect=j + 1, lx=lx, ly=ly, ux=ux, uy=uy,
                                       xs=self.w[i, j, lx - 1:ux - 1, :], ys=self.h[i, j, ly - 1:uy - 1, :]))
        return ranges

    def get(self, layer, aspect, y, x):  # Anchors.lua:60-67, 1-based indices
        w, h = self.w, self.h
        r = Rect(w[layer - 1, aspect - 1, x - 1, 0], h[layer - 1, aspect - 1, y - 1, 0],
                 w[layer - 1, aspect - 1, x - 1, 1], h[layer - 1, aspect - 1, y - 1, 1])
        r.layer, r.aspect = layer, aspect
        r.index = ((aspect * 6 - 5, aspect * 6), y, x)
        return r

    def findPositive(self, roi_list, clip_rect, pos_threshold, neg_threshold, include_best):  # Anchors.lua:147-195
        """roi_list: list of dicts with 'rect' (Rect).  Returns [(anchor_rect, roi), ...] in the reference's order; the
        nested loops over anchors x ROIs run in one CUDA launch (frcnn_find_positive), one CTA per ROI."""
        n = len(roi_list)
        if n == 0:
            return []
        rois = np.ascontiguousarray([list(r["rect"].unpack()) for r in roi_list], dtype=np.float64).reshape(-1)
        clip = ffi.new("double[4]", list(clip_rect.unpack())) if clip_rect is not None else ffi.NULL
        cap = 1 << 16
        while True:
            out, out_roi, n_out = ffi.new("frcnn_anchor_ref[]", cap), ffi.new("int[]", cap), ffi.new("int*")
            rc = lib().frcnn_find_positive(self.model.ctx, ffi.cast("const double*", rois.ctypes.data), n, clip, pos_threshold,
                                           neg_threshold, 1 if include_best else 0, out, out_roi, cap, n_out)
            if rc == 6 and cap < (1 << 22):  # FRCNN_E_OVERFLOW: more matches than the buffer
                cap *= 4
                continue
            check(self.model.ctx, rc)
            break
        return [(self.get(out[i].layer, out[i].aspect, out[i].y, out[i].x), roi_list[out_roi[i]]) for i in range(n_out[0])]

    def nearbyNegative(self, positive, neg_threshold):  # BatchIterator.lua:206-217 (before shuffle_n)
        """The nearby-aversion candidates of the positives [(anchor_rect, roi), ...]: findNearby(p:center()) filtered by
        Rect.IoU(p, a) < neg_threshold, in the reference's order, as one launch (frcnn_find_nearby_negative).  Returns
        [(anchor_rect,), ...] like the `nearby_negative` table."""
        n = len(positive)
        if n == 0:
            return []
        pos = ffi.new("frcnn_anchor_ref[]", n)
        for i, p in enumerate(positive):
            a = p[0]
            pos[i].layer, pos[i].aspect, pos[i].y, pos[i].x = a.layer, a.aspect, a.index[1], a.index[2]
        cap = 64 * n
        out, out_pos, n_out = ffi.new("frcnn_anchor_ref[]", cap), ffi.new("int[]", cap), ffi.new("int*")
        check(self.model.ctx, lib().frcnn_find_nearby_negative(self.model.ctx, pos, n, neg_threshold, out, out_pos, cap, n_out))
        return [(self.get(out[i].layer, out[i].aspect, out[i].y, out[i].x),) for i in range(n_out[0])]

    def sampleNegative(self, image_rect, roi_list, neg_threshold, count, rnd, retry=0, return_retry=False):  # Anchors.lua:197-235
        """`rnd`: numpy uint32 array of torch.random() values, three per trial (the RNG contract: the caller owns the
        generator).  Returns ([(anchor_rect,), ...], trials consumed, finished[, retry]); `retry` continues a loop whose
        random stream ran out (pass the returned retry and the remaining count)."""
        rnd = np.ascontiguousarray(rnd, dtype=np.uint32)
        n_trials = len(rnd) // 3
        rois = np.ascontiguousarray([list(r["rect"].unpack()) for r in roi_list], dtype=np.float64).reshape(-1)
        img = ffi.new("double[4]", list(image_rect.unpack()))
        cap = max(count, 1)
        out = ffi.new("frcnn_anchor_ref[]", cap)
        n_out, used, fin, rt = ffi.new("int*"), ffi.new("int*"), ffi.new("int*"), ffi.new("int*")
        check(self.model.ctx, lib().frcnn_sample_negative(self.model.ctx, img, ffi.cast("const double*", rois.ctypes.data) if len(roi_list) else ffi.NULL,
                                                          len(roi_list), neg_threshold, count, ffi.cast("const uint32_t*", rnd.ctypes.data),
                                                          n_trials, int(retry), out, cap, n_out, used, fin, rt))
        res = [(self.get(out[i].layer, out[i].aspect, out[i].y, out[i].x),) for i in range(n_out[0])]
        return (res, used[0], bool(fin[0]), rt[0]) if return_retry else (res, used[0], bool(fin[0]))

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
