"""Host mirror of Detector.lua: Detector(model), Detector:detect(input) -> list of winners."""
import numpy as np
import torch

from ._lib import check, ffi, lib
from .geometry import Anchors, Localizer
from .rect import Rect


def extract_roi_pooling_input(model, rects, fmap):
    """extract_roi_pooling_input + amp:forward (objective.lua:5-13, Detector.lua:96-97) for a list of rects on a
    Torch-layout fp32 feature map [C][H][W] (CUDA tensor).  Returns (out [R][C*kh*kw], argmax [R][C*kh*kw])."""
    fmap = fmap.to(torch.float32).contiguous()
    C, H, W = fmap.shape
    R = len(rects)
    kh, kw = model.cfg["roi_pooling"]["kh"], model.cfg["roi_pooling"]["kw"]
    out = torch.empty((R, C * kh * kw), dtype=torch.float32, device=fmap.device)
    arg = torch.empty((R, C * kh * kw), dtype=torch.int32, device=fmap.device)
    flat = np.ascontiguousarray([list(r.unpack()) for r in rects], dtype=np.float64).reshape(-1)
    check(model.ctx, lib().frcnn_roi_pool_forward(model.ctx, ffi.cast("const float*", fmap.data_ptr()), C, H, W,
                                                  ffi.cast("const double*", flat.ctypes.data), R,
                                                  ffi.cast("float*", out.data_ptr()), ffi.cast("int32_t*", arg.data_ptr())))
    return out, arg


class Detector:
    def __init__(self, model):  # Detector.lua:8-15
        self.model = model
        self.anchors = Anchors(model, model.cfg["scales"])
        self.localizer = Localizer(model, model.n_heads + 1)
        self._cap = 4096
        self._out = ffi.new("frcnn_detection[]", self._cap)

    def decode(self, head_outputs, h, w, threshold=0.95):
        """The match list of Detector.lua:36-66 from the 4 anchor-head maps (CUDA tensors [18][hi][wi])."""
        heads = [t.to(torch.float32).contiguous() for t in head_outputs]
        ptrs = ffi.new("const float*[]", [ffi.cast("const float*", t.data_ptr()) for t in heads])
        cap = sum(t.shape[-1] * t.shape[-2] * 3 for t in heads)
        buf = ffi.new("frcnn_candidate[]", cap)
        n = ffi.new("int*")
        ctx = self.model.ctx
        check(ctx, lib().frcnn_rpn_decode(ctx, ptrs, h, w, threshold, buf, cap, n))
        out = []
        for i in range(n[0]):
            c = buf[i]
            a = self.anchors.get(c.layer, c.aspect, c.y, c.x)
            out.append(dict(p=np.float32(c.logp), a=a, r=Rect(c.r[0], c.r[1], c.r[2], c.r[3]), l=c.layer,
                            box=np.array([c.box[0], c.box[1], c.box[2], c.box[3]], dtype=np.float32)))
        return out

    def _winners(self, n):
        res = []
        for i in range(n):
            d = self._out[i]
            a = self.anchors.get(d.layer, d.aspect, d.y, d.x)
            res.append({"p": np.float32(d.p), "a": a, "r": Rect(*[d.r[k] for k in range(4)]), "l": d.layer,
                        "r2": Rect(*[d.r2[k] for k in range(4)]), "class": d.cls, "confidence": np.float32(d.confidence),
                        "image": d.image})
        return res

    def detect(self, input):  # Detector.lua:17-141
        """input: [3][H][W] (or [N][3][H][W]) fp32, host (numpy / CPU tensor: copied to the GPU inside the call, like
        Detector.lua:32) or CUDA tensor.  Returns the winners {p, a, r, l, r2, class, confidence} grouped by class
        ascending, pick order inside a class (the reference's cross-class order is unspecified)."""
        ctx = self.model.ctx
        n_det = ffi.new("int*")
        if torch.is_tensor(input) and input.is_cuda:
            x = input.to(torch.float32).contiguous()
            x4 = x if x.dim() == 4 else x.unsqueeze(0)
            n, _, h, w = x4.shape
            check(ctx, lib().frcnn_detect_dev(ctx, ffi.cast("const float*", x4.data_ptr()), n, h, w, self._out, self._cap, n_det))
        else:
            x = np.ascontiguousarray(input.numpy() if torch.is_tensor(input) else input, dtype=np.float32)
            x4 = x if x.ndim == 4 else x[None]
            n, _, h, w = x4.shape
            check(ctx, lib().frcnn_detect(ctx, ffi.cast("const float*", x4.ctypes.data), n, h, w, self._out, self._cap, n_det))
        return self._winners(n_det[0])

    def stats(self):
        s = ffi.new("int64_t[4]")
        check(self.model.ctx, lib().frcnn_detect_stats(self.model.ctx, s))
        return dict(matches=int(s[0]), candidates=int(s[1]), classified=int(s[2]), winners=int(s[3]))
