"""ctypes loader for oracle/nms_ref.c (test infrastructure / CPU baseline only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_nms.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_nms.restype = ctypes.c_int64
        _LIB.oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_float, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_void_p]
        _LIB.oracle_nms_pairs.restype = ctypes.c_int64
        _LIB.oracle_nms_pairs.argtypes = [ctypes.c_int]
        _LIB.oracle_nms_segmented.restype = None
        _LIB.oracle_nms_segmented.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int,
                                              ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_int]
    return _LIB


def nms(boxes, overlap, order_mode=0, order_col=0):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    pick = np.empty((max(n, 1),), dtype=np.int64)
    c = lib().oracle_nms(boxes.ctypes.data, n, boxes.shape[1] if n else 4, overlap, order_mode, order_col,
                         pick.ctypes.data)
    return pick[:c].copy()


def nms_segmented(boxes, seg_offsets, overlap, order_mode=0, order_col=0, threads=1):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    seg = np.ascontiguousarray(seg_offsets, dtype=np.int64)
    n_seg = len(seg) - 1
    pick = np.empty((max(boxes.shape[0], 1),), dtype=np.int64)
    counts = np.zeros((n_seg,), dtype=np.int64)
    lib().oracle_nms_segmented(boxes.ctypes.data, boxes.shape[1], seg.ctypes.data, n_seg, overlap, order_mode,
                               order_col, pick.ctypes.data, counts.ctypes.data, threads)
    return pick, counts


def pair_iou_count(reset=True):
    """Pair-IoU evaluations the reference algorithm performed since the last reset (nms.lua:72-96)."""
    return int(lib().oracle_nms_pairs(1 if reset else 0))
