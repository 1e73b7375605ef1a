"""ctypes front-end of oracle/box_oracle.c (TEST INFRASTRUCTURE ONLY).

Follows nerf_rpn/model/utils.py:215-265,387-458 and nerf_rpn/model/rotated_iou/*.py; see the C file
header for the line-by-line citations and the arithmetic conventions.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libbox_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libbox_oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        fp = ctypes.POINTER(ctypes.c_float)
        L.orc_iou_matrix.argtypes = [fp, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int, fp]
        L.orc_iou_pairs.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, fp]
        L.orc_nms.argtypes = [fp, ctypes.c_int, fp, ctypes.c_int, ctypes.c_float, ctypes.POINTER(ctypes.c_int64)]
        L.orc_nms.restype = ctypes.c_int
        L.orc_batched_nms.argtypes = [fp, ctypes.c_int, fp, ctypes.POINTER(ctypes.c_int32), ctypes.c_int,
                                      ctypes.c_float, ctypes.POINTER(ctypes.c_int64)]
        L.orc_batched_nms.restype = ctypes.c_int
        L.orc_sort_vertices.argtypes = [fp, ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int32),
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def iou_matrix(a, b):
    a = _f32(a); b = _f32(b)
    assert a.ndim == 2 and b.ndim == 2 and a.shape[1] == b.shape[1] and a.shape[1] in (6, 7)
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float32)
    lib().orc_iou_matrix(_p(a, ctypes.c_float), a.shape[0], _p(b, ctypes.c_float), b.shape[0], a.shape[1],
                         _p(out, ctypes.c_float))
    return out


def iou_pairs(a, b):
    a = _f32(a); b = _f32(b)
    assert a.shape == b.shape and a.shape[1] in (6, 7)
    out = np.empty((a.shape[0],), dtype=np.float32)
    lib().orc_iou_pairs(_p(a, ctypes.c_float), _p(b, ctypes.c_float), a.shape[0], a.shape[1], _p(out, ctypes.c_float))
    return out


def nms(boxes, scores, thr):
    boxes = _f32(boxes); scores = _f32(scores)
    n = boxes.shape[0]
    keep = np.empty((max(n, 1),), dtype=np.int64)
    k = lib().orc_nms(_p(boxes, ctypes.c_float), boxes.shape[1] if n else 6, _p(scores, ctypes.c_float), n,
                      ctypes.c_float(thr), _p(keep, ctypes.c_int64))
    return keep[:k].copy()


def batched_nms(boxes, scores, groups, thr):
    boxes = _f32(boxes); scores = _f32(scores)
    groups = np.ascontiguousarray(np.asarray(groups, dtype=np.int32))
    n = boxes.shape[0]
    keep = np.empty((max(n, 1),), dtype=np.int64)
    k = lib().orc_batched_nms(_p(boxes, ctypes.c_float), boxes.shape[1] if n else 6, _p(scores, ctypes.c_float),
                              _p(groups, ctypes.c_int32), n, ctypes.c_float(thr), _p(keep, ctypes.c_int64))
    return keep[:k].copy()


def sort_vertices(vertices, mask, num_valid):
    v = _f32(vertices)
    mk = np.ascontiguousarray(np.asarray(mask, dtype=np.uint8))
    nv = np.ascontiguousarray(np.asarray(num_valid, dtype=np.int32))
    b, n, m, _ = v.shape
    idx = np.empty((b, n, 9), dtype=np.int32)
    lib().orc_sort_vertices(_p(v, ctypes.c_float), _p(mk, ctypes.c_uint8), _p(nv, ctypes.c_int32), b, n, m,
                            _p(idx, ctypes.c_int32))
    return idx
