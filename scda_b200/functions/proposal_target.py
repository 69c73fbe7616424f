"""RCNN proposal targets (mirrors functions/proposal_target.py:17-177 of the reference).

Reference: numpy on the host — append GTs, clip, cython IoU R x G, positive / negative
sets, np.random.choice sub-sampling to 128 positives + negatives up to 512, class-specific
box targets normalised by the precomputed stds, padding to 512 by resampling, then four
H2D copies.  Here the same steps run on the device with static shapes; the IoU comes from
scda_bbox_overlaps (the cython arithmetic, bit for bit).

`compute_proposal_targets` keeps the reference's signature (proposals [N, >=5] on any
device) and returns the same four CUDA tensors.  `proposal_targets_device` takes the
fixed-capacity buffer + device count produced by rpn_proposals_device and never
synchronises.
"""
import ctypes as C

import torch

from .._lib import check, load, stream_ptr
from ..extensions._cython_bbox.cython_bbox import bbox_overlaps_device
from ..utils.bbox_helper import clip_t, encode_t
from . import _sampling


import os as _os
_FORCE_TENSOR_OPS = _os.environ.get('SCDA_PT_TENSOR_OPS', '0') == '1'      # debugging switch


def proposal_targets_device(boxes, n_boxes, gts, cfg, image_hw, batch_ix=0, rng=None):
    """boxes float32 [cap, >=4] (x1,y1,x2,y2,...), n_boxes 0-dim int tensor (live rows),
    gts float32 [G, 5] (padded rows allowed).  Returns rois [bs,5], labels [bs] int64,
    loc_targets / loc_weights [bs, num_classes*4] float32 with bs = cfg['batch_size'].
    ONE kernel (csrc/target_ops.cu) for up to 4096 boxes and 256 ground truths; the tensor-op form of the
    same steps (`proposal_targets_tensor_ops`) beyond that."""
    rng = rng or _sampling.TorchRng()
    cap, G = boxes.shape[0], gts.shape[0]
    R = cap + G if cfg['append_gts'] else cap
    if not (boxes.is_cuda and R <= 4096 and 0 < G <= 256 and cfg['batch_size'] <= 4096) or _FORCE_TENSOR_OPS:
        return proposal_targets_tensor_ops(boxes, n_boxes, gts, cfg, image_hw, batch_ix, rng)
    dev = boxes.device
    bs, nc = cfg['batch_size'], cfg['num_classes']
    boxes = boxes.float()
    if boxes.stride(1) != 1:
        boxes = boxes.contiguous()
    gts = gts.float().contiguous()
    assert gts.shape[1] >= 5
    if gts.shape[1] != 5:
        gts = gts[:, :5].contiguous()
    n_dev = n_boxes.to(device=dev, dtype=torch.int64).reshape(1) if torch.is_tensor(n_boxes) else \
        torch.full((1,), int(n_boxes), dtype=torch.int64, device=dev)
    # the three draws of the reference, in its order (positives, negatives, padding)
    k_pos, k_neg, k_pad = rng.uniform(R, dev), rng.uniform(R, dev), rng.uniform(bs, dev)
    rois = torch.empty(bs, 5, dtype=torch.float32, device=dev)
    labels = torch.empty(bs, dtype=torch.int64, device=dev)
    loc_t = torch.empty(bs, nc * 4, dtype=torch.float32, device=dev)
    loc_w = torch.empty(bs, nc * 4, dtype=torch.float32, device=dev)
    norm = bool(cfg['bbox_normalize_stats_precomputed'])
    means = (C.c_double * 4)(*[float(v) for v in cfg['bbox_normalize_means']]) if norm else None
    stds = (C.c_double * 4)(*[float(v) for v in cfg['bbox_normalize_stds']]) if norm else None
    h, w = image_hw
    with torch.cuda.device(dev):
        check(load().scda_proposal_targets(
            cap, boxes.stride(0), boxes.data_ptr(), n_dev.data_ptr(), G, gts.data_ptr(),
            1 if cfg['append_gts'] else 0, float(h), float(w), float(cfg['positive_iou_thresh']),
            float(cfg['negative_iou_thresh_hi']), float(cfg['negative_iou_thresh_lo']),
            int(cfg['positive_percent'] * bs), bs, nc, 1 if norm else 0,
            C.cast(means, C.c_void_p) if norm else None, C.cast(stds, C.c_void_p) if norm else None,
            float(batch_ix), k_pos.data_ptr(), k_neg.data_ptr(), k_pad.data_ptr(), rois.data_ptr(),
            labels.data_ptr(), loc_t.data_ptr(), loc_w.data_ptr(), stream_ptr(dev)), "scda_proposal_targets")
    return rois, labels, loc_t, loc_w


def proposal_targets_tensor_ops(boxes, n_boxes, gts, cfg, image_hw, batch_ix=0, rng=None):
    """the same computation as ~150 tensor operations (shapes beyond the kernel's shared memory; also the
    second implementation the kernel is tested against)"""
    rng = rng or _sampling.TorchRng()
    dev = boxes.device
    cap, G = boxes.shape[0], gts.shape[0]
    bs, nc = cfg['batch_size'], cfg['num_classes']
    h, w = image_hw
    gt_ok = (gts[:, 2] > gts[:, 0] + 1) & (gts[:, 3] > gts[:, 1] + 1)
    live = torch.arange(cap, device=dev) < n_boxes
    if cfg['append_gts']:
        rois = torch.cat([boxes[:, :4], gts[:, :4]], dim=0)
        live = torch.cat([live, gt_ok])
    else:
        rois = boxes[:, :4]
    rois = clip_t(rois.float(), h, w).contiguous()
    R = rois.shape[0]
    ov = bbox_overlaps_device(rois, gts[:, :4].contiguous())                  # [R, G]
    ov = torch.where(gt_ok.unsqueeze(0), ov, torch.full_like(ov, -1.0))       # filtered-out GT rows
    mx, argmax = ov.max(dim=1)
    pos_m = live & (mx > cfg['positive_iou_thresh'])
    neg_m = live & (mx < cfg['negative_iou_thresh_hi']) & (mx >= cfg['negative_iou_thresh_lo']) & ~pos_m

    want_pos = int(cfg['positive_percent'] * bs)
    pos_idx, n_pos = _sampling.choose(pos_m, want_pos, rng, bs)
    neg_idx, n_neg = _sampling.choose(neg_m, bs - n_pos, rng, bs)
    total = n_pos + n_neg
    # pad to bs by resampling rows [0, total) with replacement (:149-155)
    u = rng.uniform(bs, dev)
    r = torch.arange(bs, device=dev)
    rep = (u * total.to(u.dtype)).floor().to(torch.int64).clamp(max=bs - 1)
    src = torch.where(r < total, r, rep[(r - total).clamp(min=0)])
    is_pos = src < n_pos
    roi_ix = torch.where(is_pos, pos_idx[src.clamp(max=bs - 1)],
                         neg_idx[(src - n_pos).clamp(min=0, max=bs - 1)])
    sel = rois[roi_ix]                                                        # [bs, 4]
    gt_sel = gts[argmax[roi_ix]]                                              # [bs, 5]
    labels = torch.where(is_pos, gt_sel[:, 4].to(torch.int32).to(torch.int64),
                         torch.zeros(bs, dtype=torch.int64, device=dev))
    t = encode_t(sel, gt_sel[:, :4])                                          # float32, like numpy
    t = torch.where(is_pos.unsqueeze(1), t, torch.zeros_like(t))              # keep log() of junk out
    t = t.double()
    if cfg['bbox_normalize_stats_precomputed']:
        means = _sampling.const_tensor(cfg['bbox_normalize_means'], torch.float64, dev)
        stds = _sampling.const_tensor(cfg['bbox_normalize_stds'], torch.float64, dev)
        t = (t - means) / stds
    onehot = torch.zeros(bs, nc, dtype=torch.bool, device=dev)
    onehot.scatter_(1, labels.clamp(min=0, max=nc - 1).unsqueeze(1), is_pos.unsqueeze(1))
    loc_w = onehot.unsqueeze(2).expand(-1, -1, 4).reshape(bs, nc * 4).float()
    loc_t = (onehot.unsqueeze(2) * t.unsqueeze(1)).reshape(bs, nc * 4).float()
    out_rois = torch.cat([torch.full((bs, 1), float(batch_ix), device=dev), sel], dim=1)
    return out_rois.contiguous(), labels.contiguous(), loc_t.contiguous(), loc_w.contiguous()


def compute_proposal_targets(proposals, cfg, ground_truth_bboxes, image_info, ignore_regions=None,
                             use_ohem=False, rng=None):
    '''
    :argument
        proposals:[N, k], k>=5, batch_idx, x1, y1, x2, y2
        ground_truth_bboxes: [batch, max_num_gts, k], k>=5, x1,y1,x2,y2,label
    returns:
        rois: [N, 5]  cls_targets: [N]  loc_targets, loc_weights: [N, num_classes * 4]   (CUDA)
    '''
    if ignore_regions is not None or use_ohem:
        raise NotImplementedError("ignore_regions / OHEM are not on the SCDA hot path")
    dev = ground_truth_bboxes.device if ground_truth_bboxes.is_cuda else torch.device("cuda")
    gts_all = ground_truth_bboxes.to(dev).float()
    props = proposals.to(dev).float()
    if torch.is_tensor(image_info):
        image_info = image_info.cpu().numpy()
    outs = []
    for b in range(gts_all.shape[0]):
        sel = props[props[:, 0] == b][:, 1:5].contiguous()      # API path: dynamic count is known
        n = torch.full((), sel.shape[0], dtype=torch.int64, device=dev)
        if sel.shape[0] == 0:
            sel = torch.zeros(1, 4, device=dev)
        outs.append(proposal_targets_device(sel, n, gts_all[b], cfg,
                                            (float(image_info[b][0]), float(image_info[b][1])),
                                            batch_ix=b, rng=rng))
    return tuple(torch.cat([o[i] for o in outs], dim=0).contiguous() for i in range(4))
