"""Restatement of Anchors.lua (LUTs, get, (de)parametrisation, training-label matching).  Test infrastructure only."""
import math

import numpy as np

from .localizer import Localizer, head_layer_info
from .rect import Rect

BIN_SIZE = 16  # Anchors.lua:5


class Anchors:
    def __init__(self, layers, anchor_nets, scales):  # Anchors.lua:7-58
        self.localizers = [Localizer(head_layer_info(layers, anchor_nets[i])) for i in range(len(scales))]
        width, height = 200, 200  # Anchors.lua:15
        # torch.Tensor == FloatTensor (main.lua:51): doubles are rounded to fp32 at store
        self.w = np.zeros((len(scales), 3, width, 2), dtype=np.float32)
        self.h = np.zeros((len(scales), 3, height, 2), dtype=np.float32)
        self.cx, self.cy = {}, {}

        def add(m, i, j, v, x):  # Anchors.lua:24-30 (1-based i, j, v kept as in Lua)
            m.setdefault(math.floor(x / BIN_SIZE), []).append((i, j, v))

        for i, s in enumerate(scales):
            a = s / math.sqrt(2)  # Anchors.lua:34
            aspects = [(s, s), (2 * a, a), (a, 2 * a)]  # Anchors.lua:35
            for j, b in enumerate(aspects):
                l = self.localizers[i]
                for y in range(1, height + 1):  # Anchors.lua:39-46
                    r = l.featureToInputRect(0, y - 1, 0, y)
                    cx, cy = r.center()
                    r = Rect.fromCenterWidthHeight(cx, cy, b[0], b[1])
                    self.h[i, j, y - 1, 0] = r.minY
                    self.h[i, j, y - 1, 1] = r.maxY
                    add(self.cy, i + 1, j + 1, y, cy)
                for x in range(1, width + 1):  # Anchors.lua:48-55
                    r = l.featureToInputRect(x - 1, 0, x, 0)
                    cx, cy = r.center()
                    r = Rect.fromCenterWidthHeight(cx, cy, b[0], b[1])
                    self.w[i, j, x - 1, 0] = r.minX
                    self.w[i, j, x - 1, 1] = r.maxX
                    add(self.cx, i + 1, j + 1, x, cx)

    def get(self, layer, aspect, y, x):  # Anchors.lua:60-67; all indices 1-based as in Lua
        w, h = self.w, self.h
        r = Rect(w[layer - 1, aspect - 1, x - 1, 0], h[layer - 1, aspect - 1, y - 1, 0],
                 w[layer - 1, aspect - 1, x - 1, 1], h[layer - 1, aspect - 1, y - 1, 1])
        r.layer, r.aspect = layer, aspect
        r.index = ((aspect * 6 - 5, aspect * 6), y, x)
        return r

    def findNearby(self, centerX, centerY):  # Anchors.lua:69-84
        found = []
        xl = self.cx.get(math.floor(centerX / BIN_SIZE))
        yl = self.cy.get(math.floor(centerY / BIN_SIZE))
        if xl and yl:
            for y in yl:
                for x in xl:
                    if y[0] == x[0] and y[1] == x[1]:
                        found.append(self.get(y[0], y[1], y[2], x[2]))
        return found

    def findRangesXY(self, rect, clip_rect=None):  # Anchors.lua:86-145 (returns 1-based lx..ux-1 ranges)
        def lower_bound(t, value):  # first 1-based index with t[i] >= value
            low, high = 1, len(t)
            while low <= high:
                mid = (low + high) // 2
                if t[mid - 1] >= value:
                    high = mid - 1
                else:
                    low = mid + 1
            return low

        def upper_bound(t, value):  # first 1-based index with t[i] > value
            low, high = 1, len(t)
            while low <= high:
                mid = (low + high) // 2
                if t[mid - 1] > value:
                    high = mid - 1
                else:
                    low = mid + 1
            return low

        ranges = []
        w, h = self.w, self.h
        for i in range(4):
            for j in range(3):
                lx = upper_bound(w[i, j, :, 1], rect.minX)
                ly = upper_bound(h[i, j, :, 1], rect.minY)
                ux = lower_bound(w[i, j, :, 0], rect.maxX)
                uy = lower_bound(h[i, j, :, 0], rect.maxY)
                if clip_rect is not None:
                    lx = max(lx, lower_bound(w[i, j, :, 0], clip_rect.minX))
                    ly = max(ly, lower_bound(h[i, j, :, 0], clip_rect.minY))
                    ux = min(ux, upper_bound(w[i, j, :, 1], clip_rect.maxX))
                    uy = min(uy, upper_bound(h[i, j, :, 1], clip_rect.maxY))
                if ux > lx and uy > ly:
                    ranges.append(dict(layer=i + 1, aspect=j + 1, lx=lx, ly=ly, ux=ux, uy=uy,
                                       xs=w[i, j, lx - 1:ux - 1, :], ys=h[i, j, ly - 1:uy - 1, :]))
        return ranges

    def findPositive(self, roi_list, clip_rect, pos_threshold, neg_threshold, include_best):  # Anchors.lua:147-195
        matches = []
        best_set, best_iou = None, None
        for roi in roi_list:
            if include_best:
                best_set, best_iou = [], -1
            for r in self.findRangesXY(roi["rect"], clip_rect):
                for y in range(1, r["ys"].shape[0] + 1):
                    minY, maxY = r["ys"][y - 1, 0], r["ys"][y - 1, 1]
                    for x in range(1, r["xs"].shape[0] + 1):
                        a = Rect(r["xs"][x - 1, 0], minY, r["xs"][x - 1, 1], maxY)
                        a.layer, a.aspect = r["layer"], r["aspect"]
                        a.index = ((a.aspect * 6 - 5, a.aspect * 6), r["ly"] + y - 1, r["lx"] + x - 1)
                        v = Rect.IoU(roi["rect"], a)
                        if v > pos_threshold:
                            matches.append((a, roi))
                            best_set = None
                        elif v > neg_threshold and best_set is not None and v >= best_iou:
                            if v - 0.025 > best_iou:
                                best_set = []
                            best_set.append(a)
                            best_iou = v
            if best_set and best_iou > 0:
                for v in best_set:
                    matches.append((v, roi))
        return matches

    def sampleNegative(self, image_rect, roi_list, neg_threshold, count, rnd):  # Anchors.lua:197-235
        """`rnd`: iterator over the values torch.random() returns (uint32): three per trial (range, x, y).  Returns the
        accepted anchors and the number of trials run."""
        ranges = self.findRangesXY(image_rect, image_rect)
        neg, retry, trials = [], 0, 0
        rnd = iter(rnd)
        while len(neg) < count and retry < 500:
            try:
                r1, r2, r3 = int(next(rnd)), int(next(rnd)), int(next(rnd))
            except StopIteration:
                break
            trials += 1
            r = ranges[r1 % len(ranges)]
            x = r2 % r["xs"].shape[0] + 1
            y = r3 % r["ys"].shape[0] + 1
            a = Rect(r["xs"][x - 1, 0], r["ys"][y - 1, 0], r["xs"][x - 1, 1], r["ys"][y - 1, 1])
            a.layer, a.aspect = r["layer"], r["aspect"]
            a.index = ((a.aspect * 6 - 5, a.aspect * 6), r["ly"] + y - 1, r["lx"] + x - 1)
            if any(Rect.IoU(roi["rect"], a) > neg_threshold for roi in roi_list):
                retry += 1
            else:
                retry = 0
                neg.append((a,))
        return neg, trials

    @staticmethod
    def inputToAnchor(anchor, rect):  # Anchors.lua:237-243 -> FloatTensor(4)
        x = (rect.minX - anchor.minX) / anchor.width()
        y = (rect.minY - anchor.minY) / anchor.height()
        w = math.log(rect.width() / anchor.width())
        h = math.log(rect.height() / anchor.height())
        return np.array([x, y, w, h], dtype=np.float32)

    @staticmethod
    def anchorToInput(anchor, t):  # Anchors.lua:245-252; t[i] are fp32 read as Lua doubles
        t = [float(v) for v in t]
        return Rect.fromXYWidthHeight(
            t[0] * anchor.width() + anchor.minX,
            t[1] * anchor.height() + anchor.minY,
            math.exp(t[2]) * anchor.width(),
            math.exp(t[3]) * anchor.height(),
        )
