"""Restatement of nms.lua:23-102 in numpy fp32.  Test infrastructure only.

Tie order of ``scores:sort(1)`` (TH quicksort, unstable) is PARITY UNPINNED; the oracle defines
it as a stable ascending sort by (key, index): among equal keys the highest index is picked first.
Returns 0-based indices in pick order (the reference returns 1-based LongTensor).
"""
import numpy as np

ORDER_Y2 = 0      # nms.lua:41-42: anything that is neither a number nor 'area' (incl. a score tensor, Q1)
ORDER_AREA = 1    # nms.lua:39-40
ORDER_COLUMN = 2  # nms.lua:37-38


def nms(boxes, overlap, scores=None):
    boxes = np.asarray(boxes, dtype=np.float32)
    if boxes.size == 0:  # nms.lua:26-28
        return np.zeros((0,), dtype=np.int64)
    one = np.float32(1)
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    area = ((x2 - x1) + one) * ((y2 - y1) + one)  # nms.lua:35, fp32 op by op
    if isinstance(scores, (int, np.integer)) and not isinstance(scores, bool):
        key = boxes[:, scores - 1]  # 1-based column, nms.lua:37-38
    elif isinstance(scores, str) and scores == "area":
        key = area
    else:
        key = y2  # nms.lua:41-42 (Q1)
    I = np.argsort(key, kind="stable")  # nms.lua:45 (+ tie rule above)
    thr = np.float32(overlap)  # THTensor_(leValue) takes a `real` (float) value
    pick = []
    with np.errstate(divide="ignore", invalid="ignore"):
        while I.size > 0:  # nms.lua:58-97
            i = I[-1]
            pick.append(i)
            if I.size == 1:
                break
            I = I[:-1]
            xx1 = np.maximum(x1[I], x1[i])
            yy1 = np.maximum(y1[I], y1[i])
            xx2 = np.minimum(x2[I], x2[i])
            yy2 = np.minimum(y2[I], y2[i])
            w = np.maximum((xx2 - xx1) + one, np.float32(0))  # nms.lua:85
            h = np.maximum((yy2 - yy1) + one, np.float32(0))  # nms.lua:86
            inter = w * h
            iou = inter / ((area[I] + area[i]) - inter)  # nms.lua:93-94
            I = I[iou <= thr]  # nms.lua:96 (`le`; NaN drops the box without picking it)
    return np.asarray(pick, dtype=np.int64)


def nms_segmented(boxes, seg_offsets, overlap, scores=None):
    """Per-class NMS as Detector.lua:125-136 runs it: independent nms() per segment; picks are
    segment-local 0-based indices, concatenated in segment order."""
    picks, counts = [], []
    for s in range(len(seg_offsets) - 1):
        p = nms(boxes[seg_offsets[s]:seg_offsets[s + 1]], overlap, scores)
        picks.append(p)
        counts.append(len(p))
    return (np.concatenate(picks) if picks else np.zeros((0,), np.int64)), np.asarray(counts, dtype=np.int64)
