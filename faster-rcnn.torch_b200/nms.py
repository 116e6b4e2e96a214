"""Host mirror of the global nms(boxes, overlap, scores) (nms.lua:23-102), running the CUDA kernels."""
import numpy as np

from ._lib import check, ffi, lib

ORDER_Y2, ORDER_AREA, ORDER_COLUMN = 0, 1, 2

_default_ctx = None


def _ctx(model=None):
    global _default_ctx
    if model is not None:
        return model.ctx
    if _default_ctx is None:
        p = ffi.new("frcnn_ctx**")
        check(None, lib().frcnn_create(p, 0, ffi.NULL))
        _default_ctx = p[0]
    return _default_ctx


def _order(scores):
    """nms.lua:37-43: a number selects that (1-based) column, the string 'area' the area, ANYTHING else --
    including a tensor of scores -- falls through to max-y."""
    if isinstance(scores, (int, np.integer)) and not isinstance(scores, bool):
        return ORDER_COLUMN, int(scores) - 1
    if isinstance(scores, str) and scores == "area":
        return ORDER_AREA, 0
    return ORDER_Y2, 0


def nms(boxes, overlap, scores=None, model=None):
    """boxes: [n][>=4] float32 rows {min_x, min_y, max_x, max_y, ...} (host array).  Returns the picked indices in
    pick order as int64, 0-BASED (the Lua shim adds 1 to produce the reference's LongTensor)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0] if boxes.ndim == 2 else 0
    if boxes.size == 0:
        return np.zeros((0,), dtype=np.int64)
    mode, col = _order(scores)
    pick = np.empty((n,), dtype=np.int64)
    cnt = ffi.new("int64_t*")
    ctx = _ctx(model)
    check(ctx, lib().frcnn_nms(ctx, ffi.cast("const float*", boxes.ctypes.data), n, boxes.shape[1], float(overlap), mode, col,
                               ffi.cast("int64_t*", pick.ctypes.data), cnt))
    return pick[:cnt[0]].copy()


def nms_segmented(boxes, seg_offsets, overlap, scores=None, model=None):
    """Per-class NMS of Detector.lua:125-136 in one call.  Returns (pick, counts): segment-local 0-based picks of
    segment s at pick[seg_offsets[s] : seg_offsets[s] + counts[s]]."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    seg = np.ascontiguousarray(seg_offsets, dtype=np.int64)
    n_seg = len(seg) - 1
    mode, col = _order(scores)
    pick = np.empty((max(int(seg[-1]), 1),), dtype=np.int64)
    counts = np.zeros((n_seg,), dtype=np.int64)
    ctx = _ctx(model)
    check(ctx, lib().frcnn_nms_segmented(ctx, ffi.cast("const float*", boxes.ctypes.data), boxes.shape[1] if boxes.ndim == 2 else 4,
                                         ffi.cast("const int64_t*", seg.ctypes.data), n_seg, float(overlap), mode, col,
                                         ffi.cast("int64_t*", pick.ctypes.data), ffi.cast("int64_t*", counts.ctypes.data)))
    return pick, counts


def nms_segmented_dev(boxes_dev, seg_offsets, overlap, scores=None, model=None):
    """Same with the boxes already resident on the GPU (torch CUDA float32 tensor [n][k]); returns CUDA tensors."""
    import torch
    seg = np.ascontiguousarray(seg_offsets, dtype=np.int64)
    n_seg = len(seg) - 1
    mode, col = _order(scores)
    n = boxes_dev.shape[0]
    pick = torch.empty((max(n, 1),), dtype=torch.int64, device=boxes_dev.device)
    counts = torch.zeros((n_seg,), dtype=torch.int64, device=boxes_dev.device)
    ctx = _ctx(model)
    check(ctx, lib().frcnn_nms_segmented_dev(ctx, ffi.cast("const float*", boxes_dev.data_ptr()), n, boxes_dev.shape[1],
                                             ffi.cast("const int64_t*", seg.ctypes.data), n_seg, float(overlap), mode, col,
                                             ffi.cast("int64_t*", pick.data_ptr()), ffi.cast("int64_t*", counts.data_ptr())))
    return pick, counts
