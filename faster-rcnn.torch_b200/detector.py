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


def roi_pooling_view(localizer, input_rect, fmap):
    """extract_roi_pooling_input exactly as objective.lua:5-13 returns it: the (non-contiguous) crop VIEW of the feature
    map [C][H][W] and its 1-based inclusive index table {{}, {y1, y2}, {x1, x2}}."""
    r = localizer.inputToFeatureRect(input_rect)
    C, H, W = fmap.shape
    r = r.clip(Rect(0, 0, W, H))
    y1, y2 = int(min(r.minY + 1, r.maxY)), int(r.maxY)
    x1, x2 = int(min(r.minX + 1, r.maxX)), int(r.maxX)
    return fmap[:, y1 - 1:y2, x1 - 1:x2], ((), (y1, y2), (x1, x2))


class SpatialAdaptiveMaxPooling:
    """The `amp` module slot (nn.SpatialAdaptiveMaxPooling(kw, kh), objective.lua:30, Detector.lua:14) for callers that keep
    the reference's per-ROI loop: forward on a strided [C][h][w] view, `indices` readable / assignable like the nn
    module's field (objective.lua:119,139,183), backward -> gradInput [C][h][w]."""

    def __init__(self, model, kw, kh):
        self.model, self.kw, self.kh = model, kw, kh
        self.indices = None
        self.output = None
        self.gradInput = None

    def forward(self, x):
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 3
        C, h, w = x.shape
        self.output = torch.empty((C, self.kh, self.kw), dtype=torch.float32, device=x.device)
        self.indices = torch.empty((C, self.kh, self.kw), dtype=torch.float32, device=x.device)
        sc, sh, sw = x.stride()
        check(self.model.ctx, lib().frcnn_adaptive_maxpool_forward(
            self.model.ctx, ffi.cast("const float*", x.data_ptr()), C, h, w, sc, sh, sw, self.kh, self.kw,
            ffi.cast("float*", self.output.data_ptr()), ffi.cast("float*", self.indices.data_ptr())))
        return self.output

    def backward(self, x, grad_output):
        C, h, w = x.shape
        g = grad_output.to(torch.float32).contiguous()
        idx = self.indices.contiguous()
        self.gradInput = torch.empty((C, h, w), dtype=torch.float32, device=x.device)
        check(self.model.ctx, lib().frcnn_adaptive_maxpool_backward(
            self.model.ctx, ffi.cast("const float*", g.data_ptr()), ffi.cast("const float*", idx.data_ptr()), C, h, w, self.kh,
            self.kw, ffi.cast("float*", self.gradInput.data_ptr())))
        return self.gradInput


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

    # -- the same call in two halves (frcnn_detect_begin / frcnn_detect_end): several frames in flight ----------
    def detect_begin(self, input):
        """Enqueues Detector:detect for `input` (host array / CPU tensor / CUDA tensor) and returns immediately.  The
        frame is copied into the context's staging buffer asynchronously: the detector keeps a reference to `input` until
        detect_end."""
        ctx = self.model.ctx
        if torch.is_tensor(input) and input.is_cuda:
            x = input.to(torch.float32).contiguous()
            x4 = x if x.dim() == 4 else x.unsqueeze(0)
            n, _, h, w = x4.shape
            self._keep = x4
            check(ctx, lib().frcnn_detect_begin(ctx, ffi.cast("const float*", x4.data_ptr()), 1, n, h, w))
        else:
            x = np.ascontiguousarray(input.numpy() if torch.is_tensor(input) else input, dtype=np.float32)
            x4 = x if x.ndim == 4 else x[None]
            n, _, h, w = x4.shape
            self._keep = x4
            check(ctx, lib().frcnn_detect_begin(ctx, ffi.cast("const float*", x4.ctypes.data), 0, n, h, w))

    def detect_end(self):
        ctx = self.model.ctx
        n_det = ffi.new("int*")
        check(ctx, lib().frcnn_detect_end(ctx, self._out, self._cap, n_det))
        self._keep = None
        return self._winners(n_det[0])

    def stats(self):
        s = ffi.new("int64_t[4]")
        check(self.model.ctx, lib().frcnn_detect_stats(self.model.ctx, s))
        return dict(matches=int(s[0]), candidates=int(s[1]), classified=int(s[2]), winners=int(s[3]))


class DetectorPipeline:
    """`in_flight` Detector replicas (one context = stream + workspaces + CUDA graph each, same parameters) behind a
    submit / drain interface: the host loop of main.lua:198-206 over a stream of frames with several frames in
    flight.  Winners come back in submission order and are identical to Detector:detect frame by frame."""

    def __init__(self, model, in_flight=2, schedule="throughput"):
        assert in_flight >= 1
        self.models = [model.replicate() for _ in range(in_flight)]  # `model` itself keeps its own schedule / graph
        for m in self.models:
            m.set_schedule(schedule)
        self.detectors = [Detector(m) for m in self.models]
        self._busy = [False] * in_flight
        self._next = 0

    def submit(self, frame):
        """Enqueues `frame`; returns the winners of the frame that previously occupied the slot (submitted `in_flight`
        calls ago), or None while the pipeline is filling."""
        i = self._next
        self._next = (i + 1) % len(self.detectors)
        done = self.detectors[i].detect_end() if self._busy[i] else None
        self.detectors[i].detect_begin(frame)
        self._busy[i] = True
        return done

    def drain(self):
        """Winners of every frame still in flight, oldest first."""
        out = []
        for k in range(len(self.detectors)):
            i = (self._next + k) % len(self.detectors)
            if self._busy[i]:
                out.append(self.detectors[i].detect_end())
                self._busy[i] = False
        return out

    def detect_many(self, frames):
        out = []
        for f in frames:
            r = self.submit(f)
            if r is not None:
                out.append(r)
        return out + self.drain()

    def close(self):
        for m in self.models:
            m.close()
