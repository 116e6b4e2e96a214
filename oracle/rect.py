"""Restatement of Rect.lua (value type of 4 Lua doubles, half-open [min,max)).  Test infrastructure only."""
import math

import numpy as np


class Rect:
    """Rect.lua:12-26 (numeric constructor only; the table form is broken upstream)."""

    __slots__ = ("minX", "minY", "maxX", "maxY", "layer", "aspect", "index")

    def __init__(self, minX, minY, maxX, maxY):
        self.minX = float(minX)
        self.minY = float(minY)
        self.maxX = float(maxX)
        self.maxY = float(maxY)
        self.layer = None
        self.aspect = None
        self.index = None

    @staticmethod
    def empty():  # Rect.lua:26-28
        return Rect(0, 0, 0, 0)

    @staticmethod
    def fromXYWidthHeight(x, y, width, height):  # Rect.lua:30-32
        return Rect(x, y, x + width, y + height)

    @staticmethod
    def fromCenterWidthHeight(cx, cy, width, height):  # Rect.lua:34-36
        return Rect.fromXYWidthHeight(cx - width * 0.5, cy - height * 0.5, width, height)

    def inflate(self, x, y):  # Rect.lua:45-47
        return Rect(self.minX - x, self.minY - y, self.maxX + x, self.maxY + y)

    def width(self):  # Rect.lua:53-55
        return self.maxX - self.minX

    def height(self):  # Rect.lua:57-59
        return self.maxY - self.minY

    def area(self):  # Rect.lua:61-63
        return self.width() * self.height()

    def center(self):  # Rect.lua:65-67
        return (self.minX + self.maxX) / 2, (self.minY + self.maxY) / 2

    def clip(self, c):  # Rect.lua:73-80
        return Rect(
            min(max(self.minX, c.minX), c.maxX),
            min(max(self.minY, c.minY), c.maxY),
            max(min(self.maxX, c.maxX), c.minX),
            max(min(self.maxY, c.maxY), c.minY),
        )

    def overlaps(self, o):  # Rect.lua:90-93 (strict)
        return self.minX < o.maxX and self.maxX > o.minX and self.minY < o.maxY and self.maxY > o.minY

    @staticmethod
    def intersect(a, b):  # Rect.lua:126-136
        minx = max(a.minX, b.minX)
        miny = max(a.minY, b.minY)
        maxx = min(a.maxX, b.maxX)
        maxy = min(a.maxY, b.maxY)
        if maxx >= minx and maxy >= miny:
            return Rect(minx, miny, maxx, maxy)
        return Rect.empty()

    @staticmethod
    def IoU(a, b):  # Rect.lua:138-141 (no +1, unlike nms.lua:35)
        i = Rect.intersect(a, b).area()
        return i / (a.area() + b.area() - i)

    def totensor(self):  # Rect.lua:143-145; torch.Tensor is FloatTensor under main.lua:51
        return np.array([self.minX, self.minY, self.maxX, self.maxY], dtype=np.float32)

    def snapToInt(self):  # Rect.lua:147-149
        return Rect(math.floor(self.minX), math.floor(self.minY), math.ceil(self.maxX), math.ceil(self.maxY))

    def offset(self, x, y):  # Rect.lua:151-153
        return Rect(self.minX + x, self.minY + y, self.maxX + x, self.maxY + y)

    def unpack(self):  # Rect.lua:114-116
        return self.minX, self.minY, self.maxX, self.maxY

    def __repr__(self):
        return "Rect(%r, %r, %r, %r)" % self.unpack()
