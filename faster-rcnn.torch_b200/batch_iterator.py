"""Host mirror of the device-facing half of BatchIterator.lua: what happens to a frame between `image.load` and
`pnet:forward` (BatchIterator.lua:101-164, objective.lua:66), as an asynchronous pipeline on a side stream.

The reference resizes (`image.scale` to find_target_size), normalises (centering, scaling, contrastive normalisation of
the luminance channel) on the CPU and uploads with a synchronous `:cuda()` inside the training step.  Here the frame goes
up from page-locked host memory with an asynchronous copy and is resized / normalised on the GPU (frcnn_scale_frame,
frcnn_normalize_frame) on the prefetcher's own stream, `depth` frames ahead of the consumer; `get()` makes the consumer's
stream wait on the frame's event, so the upload and the preprocessing of step i + 1 overlap the compute of step i.
Disk I/O, augmentation and the epoch bookkeeping of BatchIterator.lua stay out of scope."""
import torch

from ._lib import check, ffi, lib


class FramePrefetcher:
    def __init__(self, device=0, depth=2, target_smaller_side=None, max_pixel_size=None, normalization=None):
        """normalization: None (frames arrive normalised) or dict(rgb2yuv, centering, scaling, contrastive_width) -- the
        cfg.normalization / color_space switches of config/*.lua; target_smaller_side / max_pixel_size: resize to
        find_target_size (utilities.lua:188-204) when given."""
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(device=self.device)
        pctx = ffi.new("frcnn_ctx**")
        check(None, lib().frcnn_create(pctx, device, ffi.cast("void*", self.stream.cuda_stream)))
        self.ctx = pctx[0]
        self.depth = max(1, depth)
        self.target = (target_smaller_side, max_pixel_size) if target_smaller_side else None
        self.norm = normalization
        self._queue = []      # (device tensor, event, staging tensor kept alive)
        self._free = {}       # shape -> reusable device buffers

    def close(self):
        if self.ctx is not None:
            self.stream.synchronize()
            lib().frcnn_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _buffer(self, shape):
        pool = self._free.setdefault(tuple(shape), [])
        if pool:
            t, ev = pool.pop()
            if ev is not None:
                self.stream.wait_event(ev)     # the consumer's last use of the buffer precedes its refill
            return t
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def recycle(self, t):
        """Hands a frame returned by get() back for reuse (optional).  Work already enqueued on the CURRENT stream may
        still read it: the refill waits for an event recorded here."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free.setdefault(tuple(t.shape), []).append((t, ev))

    def submit(self, host_frame):
        """host_frame: [3][h][w] fp32 CPU tensor (page-locked for a truly asynchronous copy), or a stack [n][3][h][w] of
        frames that need no resize / normalisation (one copy for the whole stack).  Returns immediately."""
        assert host_frame.dtype == torch.float32 and host_frame.dim() in (3, 4)
        L = lib()
        if host_frame.dim() == 4:
            assert not self.target and not self.norm, "stacks are uploaded as they are"
            with torch.cuda.stream(self.stream):
                raw = self._buffer(host_frame.shape)
                raw.copy_(host_frame, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self._queue.append((raw, ev, None, host_frame))
            return
        with torch.cuda.stream(self.stream):
            raw = self._buffer(host_frame.shape)
            raw.copy_(host_frame, non_blocking=True)
            out = raw
            if self.target:
                w, h = ffi.new("int*"), ffi.new("int*")
                check(None, L.frcnn_find_target_size(host_frame.shape[2], host_frame.shape[1], float(self.target[0]), float(self.target[1]), w, h))
                if (h[0], w[0]) != tuple(host_frame.shape[1:]):
                    out = self._buffer((3, h[0], w[0]))
                    check(self.ctx, L.frcnn_scale_frame(self.ctx, ffi.cast("const float*", raw.data_ptr()), 3, raw.shape[1], raw.shape[2],
                                                        ffi.cast("float*", out.data_ptr()), h[0], w[0]))
            if self.norm:
                n = self.norm
                check(self.ctx, L.frcnn_normalize_frame(self.ctx, ffi.cast("float*", out.data_ptr()), out.shape[1], out.shape[2],
                                                        1 if n.get("rgb2yuv") else 0, 1 if n.get("centering", True) else 0,
                                                        1 if n.get("scaling", True) else 0, int(n.get("contrastive_width", 7))))
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._queue.append((out, ev, raw if raw is not out else None, host_frame))

    def pending(self):
        return len(self._queue)

    def get(self):
        """The oldest submitted frame as a device tensor; the CURRENT stream waits for its upload / preprocessing."""
        out, ev, raw, _ = self._queue.pop(0)
        torch.cuda.current_stream(self.device).wait_event(ev)
        if raw is not None:
            self._free.setdefault(tuple(raw.shape), []).append((raw, None))   # only this stream ever touched it
        return out
