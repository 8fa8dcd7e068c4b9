"""ORACLE (test infrastructure) — ctypes front-end of oracle/rotated_ops.cpp.

Restates detectron2 ``nms_rotated`` / ``pairwise_iou_rotated`` (reference call sites
lib/general.py:177 and test.py:135).  PARITY UNPINNED (see oracle/__init__.py).
"""
import ctypes

import numpy as np
import torch

from . import build

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        f32p = ctypes.POINTER(ctypes.c_float)
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.oracle_pairwise_iou_rotated.argtypes = [f32p, ctypes.c_int64, f32p, ctypes.c_int64, f32p]
        lib.oracle_pairwise_iou_rotated.restype = None
        lib.oracle_nms_rotated.argtypes = [f32p, f32p, ctypes.c_int64, ctypes.c_float, ctypes.c_int, i64p]
        lib.oracle_nms_rotated.restype = ctypes.c_int64
        lib.oracle_nms_near_threshold.argtypes = [f32p, ctypes.c_int64, ctypes.c_float, ctypes.c_float]
        lib.oracle_nms_near_threshold.restype = ctypes.c_int64
        _lib = lib
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def pairwise_iou_rotated(boxes1, boxes2):
    """[N,5] x [M,5] (cx,cy,w,h,deg) -> [N,M] fp32 skew IoU."""
    a, ap = _f32(boxes1)
    b, bp = _f32(boxes2)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    if out.size:
        _load().oracle_pairwise_iou_rotated(ap, a.shape[0], bp, b.shape[0],
                                            out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return torch.from_numpy(out)


def nms_rotated(boxes, scores, iou_threshold, strict=True):
    """Greedy rotated NMS; returns int64 keep indices in descending-score (stable) order."""
    b, bp = _f32(boxes)
    s, sp = _f32(scores)
    n = b.shape[0]
    keep = np.zeros((max(n, 1),), dtype=np.int64)
    k = 0
    if n:
        k = _load().oracle_nms_rotated(bp, sp, n, float(iou_threshold), 1 if strict else 0,
                                       keep.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return torch.from_numpy(keep[:k].copy())


def near_threshold_pairs(boxes, thr, tol=1e-6):
    b, bp = _f32(boxes)
    return int(_load().oracle_nms_near_threshold(bp, b.shape[0], float(thr), float(tol)))
