"""CPU oracle for the SCDA operator hot path — numpy-facing binding.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from scda_b200/.

Each function takes and returns numpy arrays and forwards to the C restatement
in scda_oracle.c (which cites the reference file:line it follows).  Host-side
plumbing of the reference (anchors, box codec, proposal/target assignment) is
restated in numpy in oracle/host.py.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f = np.float32
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)
_up = C.POINTER(C.c_uint64)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            from . import build as _b
            _b.build_oracle()
        _LIB = C.CDLL(path)
        _LIB.oracle_nms_scan.restype = C.c_long
        _LIB.oracle_nms.restype = C.c_long
        _LIB.oracle_cpu_nms.restype = C.c_long
    return _LIB


def set_threads(n: int) -> int:
    """thread count of the C restatement's OpenMP loops; returns what is in force"""
    return int(lib().oracle_set_num_threads(int(n)))


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _p(a, t):
    return a.ctypes.data_as(t)


# ---------------------------------------------------------------- RoI pooling
def roi_pool_forward(feat, rois, ph, pw, scale, want_argmax=True):
    feat, rois = _c(feat, _f), _c(rois, _f)
    B, Ch, H, W = feat.shape
    R = rois.shape[0]
    out = np.zeros((R, Ch, ph, pw), _f)
    arg = np.zeros((R, Ch, ph, pw), np.int32) if want_argmax else None
    lib().oracle_roi_pool_forward(_p(feat, _fp), C.c_float(scale), R, H, W, Ch, ph, pw,
                                  _p(rois, _fp), _p(out, _fp),
                                  _p(arg, _ip) if want_argmax else None)
    return (out, arg) if want_argmax else out


def roi_pool_backward(top_diff, rois, argmax, feat_shape, scale):
    top_diff, rois, argmax = _c(top_diff, _f), _c(rois, _f), _c(argmax, np.int32)
    B, Ch, H, W = feat_shape
    R, _, ph, pw = top_diff.shape
    g = np.zeros(feat_shape, _f)
    lib().oracle_roi_pool_backward(_p(top_diff, _fp), C.c_float(scale), B, R, H, W, Ch, ph, pw,
                                   _p(rois, _fp), _p(g, _fp), _p(argmax, _ip))
    return g


# ------------------------------------------------------------------ RoIAlign
def roi_align_forward(feat, rois, ah, aw, scale):
    feat, rois = _c(feat, _f), _c(rois, _f)
    B, Ch, H, W = feat.shape
    R = rois.shape[0]
    out = np.zeros((R, Ch, ah, aw), _f)
    lib().oracle_roi_align_forward(_p(feat, _fp), C.c_float(scale), R, H, W, Ch, ah, aw,
                                   _p(rois, _fp), _p(out, _fp))
    return out


def roi_align_backward(top_diff, rois, feat_shape, scale):
    top_diff, rois = _c(top_diff, _f), _c(rois, _f)
    B, Ch, H, W = feat_shape
    R, _, ah, aw = top_diff.shape
    g = np.zeros(feat_shape, _f)
    lib().oracle_roi_align_backward(_p(top_diff, _fp), C.c_float(scale), B, R, H, W, Ch, ah, aw,
                                    _p(rois, _fp), _p(g, _fp))
    return g


# ----------------------------------------------------------------------- NMS
def nms_mask(boxes5, thresh):
    boxes5 = _c(boxes5, _f)
    n = boxes5.shape[0]
    cb = (n + 63) // 64
    mask = np.zeros((n, cb), np.uint64)
    lib().oracle_nms_mask(n, _p(boxes5, _fp), C.c_float(thresh), _p(mask, _up))
    return mask


def nms_scan(mask):
    mask = _c(mask, np.uint64)
    n = mask.shape[0]
    keep = np.zeros(max(n, 1), np.int64)
    k = lib().oracle_nms_scan(n, _p(mask, _up), _p(keep, _lp))
    return keep[:k].copy()


def nms(boxes5, thresh):
    """gpu_nms semantics: pre-sorted boxes, +1 IoU, suppress iff IoU > thresh."""
    boxes5 = _c(boxes5, _f)
    n = boxes5.shape[0]
    keep = np.zeros(max(n, 1), np.int64)
    k = lib().oracle_nms(n, _p(boxes5, _fp), C.c_float(thresh), _p(keep, _lp))
    return keep[:k].copy()


def cpu_nms(boxes5, thresh):
    """cpu_nms semantics (>=), ordering by descending score done here as pth_nms would."""
    boxes5 = _c(boxes5, _f)
    n = boxes5.shape[0]
    order = _c(np.argsort(-boxes5[:, 4], kind="stable"), np.int64)
    areas = _c((boxes5[:, 2] - boxes5[:, 0] + 1) * (boxes5[:, 3] - boxes5[:, 1] + 1), _f)
    keep = np.zeros(max(n, 1), np.int64)
    k = lib().oracle_cpu_nms(n, _p(boxes5, _fp), _p(order, _lp), _p(areas, _fp),
                             C.c_float(thresh), _p(keep, _lp))
    return keep[:k].copy()


# ----------------------------------------------------------------------- IoU
def bbox_overlaps(boxes, query):
    """cython_bbox.bbox_overlaps convention."""
    boxes, query = _c(boxes, _f), _c(query, _f)
    out = np.zeros((boxes.shape[0], query.shape[0]), _f)
    lib().oracle_bbox_overlaps(boxes.shape[0], _p(boxes, _fp), query.shape[0], _p(query, _fp),
                               _p(out, _fp))
    return out


def iou_overlap(b1, b2):
    """IOUOverlapKernel convention (union clamped to >= 1)."""
    b1, b2 = _c(b1[:, :4], _f), _c(b2[:, :4], _f)
    out = np.zeros((b1.shape[0], b2.shape[0]), _f)
    lib().oracle_iou_overlap(_p(b1, _fp), _p(b2, _fp), 4, b1.shape[0], b2.shape[0], _p(out, _fp))
    return out


# --------------------------------------------------------------- focal losses
def sigmoid_focal_forward(logits, targets, weight_pos, gamma, alpha):
    logits, targets = _c(logits, _f), _c(targets, np.int32)
    M, K = logits.shape
    losses = np.zeros((M, K), _f)
    lib().oracle_sigmoid_focal_forward(M * K, _p(logits, _fp), _p(targets, _ip),
                                       C.c_float(weight_pos), C.c_float(gamma), C.c_float(alpha),
                                       K, _p(losses, _fp))
    return losses


def sigmoid_focal_backward(logits, targets, weight_pos, gamma, alpha):
    logits, targets = _c(logits, _f), _c(targets, np.int32)
    M, K = logits.shape
    dx = np.zeros((M, K), _f)
    lib().oracle_sigmoid_focal_backward(M * K, _p(logits, _fp), _p(targets, _ip), _p(dx, _fp),
                                        C.c_float(weight_pos), C.c_float(gamma),
                                        C.c_float(alpha), K)
    return dx


def softmax_focal_forward(logits, targets, weight_pos, gamma, alpha):
    logits, targets = _c(logits, _f), _c(targets, np.int32)
    M, K = logits.shape
    losses, priors = np.zeros(M, _f), np.zeros((M, K), _f)
    lib().oracle_softmax_focal_forward(M * K, _p(logits, _fp), _p(targets, _ip),
                                       C.c_float(weight_pos), C.c_float(gamma), C.c_float(alpha),
                                       K, _p(losses, _fp), _p(priors, _fp))
    return losses, priors


def softmax_focal_backward(logits, targets, priors, weight_pos, gamma, alpha):
    logits, targets, priors = _c(logits, _f), _c(targets, np.int32), _c(priors, _f)
    M, K = logits.shape
    dx, buff = np.zeros((M, K), _f), np.zeros(M, _f)
    lib().oracle_softmax_focal_backward(M * K, _p(logits, _fp), _p(targets, _ip), _p(dx, _fp),
                                        C.c_float(weight_pos), C.c_float(gamma),
                                        C.c_float(alpha), K, _p(priors, _fp), _p(buff, _fp))
    return dx, buff
