"""RPN anchor targets (mirrors functions/anchor_target.py:16-116 of the reference).

The reference does this in numpy on the host: D2H of the ground truth, cython IoU of
30 720 anchors x G boxes on one core, argmax / where / np.random.choice, then three H2D
copies.  Here the IoU matrix comes from scda_bbox_overlaps (bit-identical to the cython
arithmetic) and the labelling, sub-sampling and box encoding run as device tensor ops with
static shapes and no host synchronisation.

Same signature and outputs as the reference, except that `loc_normalizer` is returned as a
0-dim device tensor instead of a Python int (it is only ever used as a divisor,
models/faster_rcnn/faster_rcnn_adver_expansion_reweight_cluster.py:54-55).
"""
import os

import torch

from .._lib import check, load, stream_ptr
from ..extensions._cython_bbox.cython_bbox import bbox_overlaps_device
from ..utils import anchor_helper
from ..utils.bbox_helper import encode_t
from . import _sampling


_FORCE_TENSOR_OPS = os.environ.get('SCDA_AT_TENSOR_OPS', '0') == '1'      # debugging switch
_A32 = {}


def _anchors32(anchors64):
    """float32 copy of the (cached) float64 anchor table, made once per table"""
    key = (anchors64.data_ptr(), tuple(anchors64.shape))
    ent = _A32.get(key)
    if ent is None or ent[0] is not anchors64:
        ent = (anchors64, anchors64.float().contiguous())
        _A32[key] = ent
    return ent[1]


def compute_anchor_targets(feature_size, cfg, ground_truth_bboxes, image_info,
                           ignore_regions=None, rng=None):
    r'''
    :argument
        feature_size: [4]. i.e. batch, num_anchors * 4, height, width
        ground_truth_bboxes: FloatTensor, [batch, max_num_gt_bboxes, 5]
        image_info: FloatTensor, [batch, 3]
        ignore_regions: must be None (the driver passes None, tools/faster_rcnn_train_val.py:520)
    :returns
        cls_targets: LongTensor [batch, num_anchors, height, width]  (-1 ignore, 0 bg, 1 fg)
        loc_targets, loc_masks: FloatTensor [batch, num_anchors * 4, height, width]
        loc_normalizer: number of anchors with label >= 0 (at least 1)
    '''
    if ignore_regions is not None:
        raise NotImplementedError("ignore_regions is not on the SCDA hot path")
    rng = rng or _sampling.TorchRng()
    B, A4, fh, fw = [int(v) for v in feature_size]
    A = A4 // 4
    assert A * 4 == A4
    KA = fh * fw * A
    dev = ground_truth_bboxes.device
    assert dev.type == "cuda", "ground_truth_bboxes must be on the GPU (no CPU path)"
    gts = ground_truth_bboxes.float().contiguous()
    anchors64 = anchor_helper.anchors_device(fh, fw, cfg['anchor_ratios'], cfg['anchor_scales'],
                                             cfg['anchor_stride'], dev)
    anchors32 = _anchors32(anchors64)
    G = gts.shape[1]
    if B == 1 and 0 < G <= 256 and KA <= 100000 and gts.shape[2] >= 5 and not _FORCE_TENSOR_OPS:
        # one kernel (csrc/target_ops.cu); the two key vectors are the draws of the reference, in its order
        k_pos, k_neg = rng.uniform(KA, dev), rng.uniform(KA, dev)
        g5 = gts[0] if gts.shape[2] == 5 else gts[0, :, :5].contiguous()
        cls_targets = torch.empty(1, A, fh, fw, dtype=torch.int64, device=dev)
        loc_targets = torch.empty(1, A * 4, fh, fw, dtype=torch.float32, device=dev)
        loc_masks = torch.empty(1, A * 4, fh, fw, dtype=torch.float32, device=dev)
        normalizer = torch.empty((), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            check(load().scda_anchor_targets(
                A, fh, fw, anchors32.data_ptr(), anchors64.data_ptr(), G, g5.data_ptr(),
                float(cfg['negative_iou_thresh']), float(cfg['positive_iou_thresh']),
                int(cfg['positive_percent'] * cfg['rpn_batch_size']), int(cfg['rpn_batch_size']),
                k_pos.data_ptr(), k_neg.data_ptr(), cls_targets.data_ptr(), loc_targets.data_ptr(),
                loc_masks.data_ptr(), normalizer.data_ptr(), stream_ptr(dev)), "scda_anchor_targets")
        return cls_targets, loc_targets, loc_masks, normalizer

    labels_all, argmax_all = [], []
    for b in range(B):
        ov = bbox_overlaps_device(anchors32, gts[b, :, :4].contiguous())       # [KA, G]
        mx, argmax = ov.max(dim=1)
        gt_max = ov.max(dim=0)[0]
        gt_max = torch.where(gt_max < 0.1, torch.full_like(gt_max, -1.0), gt_max)
        hit = ov == gt_max.unsqueeze(0)                                        # [KA, G]
        any_hit = hit.any(dim=1)
        # argmax_overlaps[gb, gka] = gg: duplicates resolve to the LAST write = largest g
        G = ov.shape[1]
        last_g = (hit.to(torch.int64) * torch.arange(1, G + 1, device=dev)).max(dim=1)[0] - 1
        argmax = torch.where(any_hit, last_g, argmax)
        lab = torch.full((KA,), -1, dtype=torch.int64, device=dev)
        lab = torch.where(mx < cfg['negative_iou_thresh'], torch.zeros_like(lab), lab)
        lab = torch.where(any_hit, torch.ones_like(lab), lab)
        lab = torch.where(mx > cfg['positive_iou_thresh'], torch.ones_like(lab), lab)
        labels_all.append(lab)
        argmax_all.append(argmax)
    labels = torch.stack(labels_all).reshape(-1)          # [B*KA], batch-major like np.where
    argmax = torch.stack(argmax_all).reshape(-1)

    n_pos_want = int(cfg['positive_percent'] * cfg['rpn_batch_size'] * B)
    pos, n_pos = _sampling.drop(labels > 0, n_pos_want, rng)
    labels = torch.where((labels > 0) & ~pos, torch.full_like(labels, -1), labels)
    n_neg_want = cfg['rpn_batch_size'] * B - n_pos
    neg, _ = _sampling.drop(labels == 0, n_neg_want, rng)
    labels = torch.where((labels == 0) & ~neg, torch.full_like(labels, -1), labels)

    pos = (labels > 0)
    b_of = torch.arange(B, device=dev).repeat_interleave(KA)
    tgt_gt = gts[b_of, argmax]                                               # [B*KA, 5] float32
    enc = encode_t(anchors64.repeat(B, 1), tgt_gt[:, :4])                     # float64
    loc_t = torch.where(pos.unsqueeze(1), enc, torch.zeros_like(enc)).float()
    loc_m = pos.unsqueeze(1).expand(-1, 4).float()

    cls_targets = labels.view(B, fh, fw, A).permute(0, 3, 1, 2).contiguous()
    loc_targets = loc_t.view(B, fh, fw, A * 4).permute(0, 3, 1, 2).contiguous()
    loc_masks = loc_m.reshape(B, fh, fw, A * 4).permute(0, 3, 1, 2).contiguous()
    loc_normalizer = (labels >= 0).sum().clamp(min=1)
    return cls_targets, loc_targets, loc_masks, loc_normalizer
