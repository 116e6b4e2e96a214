"""Host mirror of the reference's Rect value type (Rect.lua): four doubles, half-open [min, max)."""
import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np


@dataclass
class Rect:
    minX: float
    minY: float
    maxX: float
    maxY: float
    # attached by Anchors:get (Anchors.lua:63-65)
    layer: Optional[int] = field(default=None, compare=False)
    aspect: Optional[int] = field(default=None, compare=False)
    index: Optional[Tuple] = field(default=None, compare=False)

    def __post_init__(self):
        self.minX, self.minY, self.maxX, self.maxY = float(self.minX), float(self.minY), float(self.maxX), float(self.maxY)

    @staticmethod
    def new(minX, minY, maxX, maxY):
        return Rect(minX, minY, maxX, maxY)

    @staticmethod
    def empty():
        return Rect(0, 0, 0, 0)

    @staticmethod
    def fromXYWidthHeight(x, y, width, height):
        return Rect(x, y, x + width, y + height)

    @staticmethod
    def fromCenterWidthHeight(cx, cy, width, height):
        return Rect.fromXYWidthHeight(cx - width * 0.5, cy - height * 0.5, width, height)

    def scale(self, fx, fy=None):
        fy = fx if fy is None else fy
        return Rect(self.minX * fx, self.minY * fy, self.maxX * fx, self.maxY * fy)

    def inflate(self, x, y):
        return Rect(self.minX - x, self.minY - y, self.maxX + x, self.maxY + y)

    def width(self):
        return self.maxX - self.minX

    def height(self):
        return self.maxY - self.minY

    def size(self):
        return self.width(), self.height()

    def area(self):
        return self.width() * self.height()

    def center(self):
        return (self.minX + self.maxX) / 2, (self.minY + self.maxY) / 2

    def isEmpty(self):
        return self.minX == self.maxX and self.minY == self.maxY

    def clip(self, c):
        return Rect(min(max(self.minX, c.minX), c.maxX), min(max(self.minY, c.minY), c.maxY),
                    max(min(self.maxX, c.maxX), c.minX), max(min(self.maxY, c.maxY), c.minY))

    def containsPt(self, x, y):
        return self.minX <= x < self.maxX and self.minY <= y < self.maxY

    def contains(self, o):
        return self.containsPt(o.minX, o.minY) and self.containsPt(o.maxX, o.maxY)

    def overlaps(self, o):
        return self.minX < o.maxX and self.maxX > o.minX and self.minY < o.maxY and self.maxY > o.minY

    def normalize(self):
        l, r = (self.minX, self.maxX) if self.minX <= self.maxX else (self.maxX, self.minX)
        t, b = (self.minY, self.maxY) if self.minY <= self.maxY else (self.maxY, self.minY)
        return Rect(l, t, r, b)

    def unpack(self):
        return self.minX, self.minY, self.maxX, self.maxY

    @staticmethod
    def union(a, b):
        return Rect(min(a.minX, b.minX), min(a.minY, b.minY), max(a.maxX, b.maxX), max(a.maxY, b.maxY))

    @staticmethod
    def intersect(a, b):
        minx, miny = max(a.minX, b.minX), max(a.minY, b.minY)
        maxx, maxy = min(a.maxX, b.maxX), min(a.maxY, b.maxY)
        return Rect(minx, miny, maxx, maxy) if (maxx >= minx and maxy >= miny) else Rect.empty()

    @staticmethod
    def IoU(a, b):
        i = Rect.intersect(a, b).area()
        return i / (a.area() + b.area() - i)

    def totensor(self):
        """fp32, as torch.Tensor under main.lua:51's default tensor type."""
        return np.array([self.minX, self.minY, self.maxX, self.maxY], dtype=np.float32)

    def snapToInt(self):
        return Rect(math.floor(self.minX), math.floor(self.minY), math.ceil(self.maxX), math.ceil(self.maxY))

    def offset(self, x, y):
        return Rect(self.minX + x, self.minY + y, self.maxX + x, self.maxY + y)

    def vertices(self):
        return np.array([[self.minX, self.minY], [self.maxX, self.minY], [self.maxX, self.maxY], [self.minX, self.maxY]],
                        dtype=np.float32)

    def clone(self):
        return Rect(self.minX, self.minY, self.maxX, self.maxY)

    def __str__(self):
        return "{ min: (%.2f, %.2f), max: (%.2f, %.2f), size: (%.2f x %.2f) }" % (
            self.minX, self.minY, self.maxX, self.maxY, self.width(), self.height())
