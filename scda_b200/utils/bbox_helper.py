"""Box helpers (mirrors utils/bbox_helper.py of the reference).

The numpy functions keep the reference's names, argument meaning and dtypes; the `_t`
functions are the same arithmetic on CUDA tensors for the on-device pipeline
(float64 where the reference's numpy promotes to float64, float32 where it does not).
"""
import warnings

import numpy as np

from ..extensions._cython_bbox import cython_bbox


def bbox_iou_overlaps(b1, b2):
    """utils/bbox_helper.py:8-9: IoU through cython_bbox on float32 casts."""
    return cython_bbox.bbox_overlaps(np.ascontiguousarray(b1[:, :4].astype(np.float32)),
                                     np.ascontiguousarray(b2[:, :4].astype(np.float32)))


def bbox_iof_overlaps(b1, b2):
    """utils/bbox_helper.py:28-43: intersection over the FIRST box's area (area clamped to >= 1)."""
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    ix = np.maximum(np.minimum(b1[:, 2:3], b2[:, 2][None]) - np.maximum(b1[:, 0:1], b2[:, 0][None]), 0)
    iy = np.maximum(np.minimum(b1[:, 3:4], b2[:, 3][None]) - np.maximum(b1[:, 1:2], b2[:, 1][None]), 0)
    return ix * iy / np.maximum(area1[:, np.newaxis], 1)


def center_to_corner(boxes):
    return np.stack([boxes[:, 0] - boxes[:, 2] / 2., boxes[:, 1] - boxes[:, 3] / 2.,
                     boxes[:, 0] + boxes[:, 2] / 2., boxes[:, 1] + boxes[:, 3] / 2.], axis=1)


def corner_to_center(boxes):
    return np.stack([(boxes[:, 0] + boxes[:, 2]) / 2., (boxes[:, 1] + boxes[:, 3]) / 2.,
                     boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]], axis=1)


def compute_loc_targets(raw_bboxes, gt_bboxes):
    """(dx, dy, log dw, log dh) with no-+1 widths, utils/bbox_helper.py:60-76."""
    bb, gt = corner_to_center(raw_bboxes), corner_to_center(gt_bboxes)
    assert np.all(bb[:, 2] > 0) and np.all(bb[:, 3] > 0)
    return np.stack([(gt[:, 0] - bb[:, 0]) / bb[:, 2], (gt[:, 1] - bb[:, 1]) / bb[:, 3],
                     np.log(gt[:, 2] / bb[:, 2]), np.log(gt[:, 3] / bb[:, 3])], axis=1)


def compute_loc_bboxes(raw_bboxes, deltas):
    """Inverse of compute_loc_targets, utils/bbox_helper.py:79-96."""
    with warnings.catch_warnings(record=True):
        warnings.simplefilter("always")
        bb = corner_to_center(raw_bboxes)
        dt = np.stack([deltas[:, 0] * bb[:, 2] + bb[:, 0], deltas[:, 1] * bb[:, 3] + bb[:, 1],
                       np.exp(deltas[:, 2]) * bb[:, 2], np.exp(deltas[:, 3]) * bb[:, 3]], axis=1)
        return center_to_corner(dt)


def clip_bbox(bbox, img_size):
    h, w = img_size[:2]
    bbox[:, 0] = np.clip(bbox[:, 0], 0, w - 1)
    bbox[:, 1] = np.clip(bbox[:, 1], 0, h - 1)
    bbox[:, 2] = np.clip(bbox[:, 2], 0, w - 1)
    bbox[:, 3] = np.clip(bbox[:, 3], 0, h - 1)
    return bbox


def compute_recall(box_pred, box_gt):
    n_gt = box_gt.shape[0]
    if box_pred.size == 0 or n_gt == 0:
        return 0, n_gt
    ov = bbox_iou_overlaps(box_gt, box_pred)
    return int((np.max(ov, axis=1) > 0.5).sum()), n_gt


# ------------------------------------------------------------------ tensors
def decode_t(raw, deltas):
    """compute_loc_bboxes on tensors: raw [N,4] (float64 anchors or float32 rois), deltas
    float32 [N,4].  Products are formed in raw's dtype after promotion, exp stays float32 —
    the dtypes numpy gives the reference."""
    import torch
    w = raw[:, 2] - raw[:, 0]
    h = raw[:, 3] - raw[:, 1]
    cx = (raw[:, 0] + raw[:, 2]) / 2.
    cy = (raw[:, 1] + raw[:, 3]) / 2.
    dt = torch.promote_types(raw.dtype, deltas.dtype)
    d = deltas.to(dt)
    ncx = d[:, 0] * w + cx
    ncy = d[:, 1] * h + cy
    nw = torch.exp(deltas[:, 2]).to(dt) * w
    nh = torch.exp(deltas[:, 3]).to(dt) * h
    return torch.stack([ncx - nw / 2., ncy - nh / 2., ncx + nw / 2., ncy + nh / 2.], dim=1)


def encode_t(raw, gt):
    """compute_loc_targets on tensors (dtype = promotion of the two inputs)."""
    import torch
    dt = torch.promote_types(raw.dtype, gt.dtype)
    bw, bh = raw[:, 2] - raw[:, 0], raw[:, 3] - raw[:, 1]
    bx, by = (raw[:, 0] + raw[:, 2]) / 2., (raw[:, 1] + raw[:, 3]) / 2.
    gw, gh = gt[:, 2] - gt[:, 0], gt[:, 3] - gt[:, 1]
    gx, gy = (gt[:, 0] + gt[:, 2]) / 2., (gt[:, 1] + gt[:, 3]) / 2.
    return torch.stack([(gx.to(dt) - bx.to(dt)) / bw.to(dt), (gy.to(dt) - by.to(dt)) / bh.to(dt),
                        torch.log(gw.to(dt) / bw.to(dt)), torch.log(gh.to(dt) / bh.to(dt))], dim=1)


def clip_t(boxes, height, width):
    import torch
    return torch.stack([boxes[:, 0].clamp(0, width - 1), boxes[:, 1].clamp(0, height - 1),
                        boxes[:, 2].clamp(0, width - 1), boxes[:, 3].clamp(0, height - 1)], dim=1)
