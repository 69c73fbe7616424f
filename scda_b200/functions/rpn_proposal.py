"""RPN proposal generation (mirrors functions/rpn_proposal.py:17-74 of the reference).

Reference flow per image: D2H of the class and delta maps -> numpy argpartition/argsort
top-k -> float64 decode -> clip -> min-size filter -> H2D -> GPU bitmask NMS -> D2H of
the 18 MB mask -> host scan -> slice -> numpy stack -> CPU tensor.
Here everything up to the final slice stays on the device and nothing synchronises:
scda_rpn_proposal_rows (csrc/proposal_ops.cu: radix selection of the pre_nms_top_n best scores,
their descending order by rank counting, decode/clip in float64 — the dtype numpy gives the
reference —, min-size filter, stable compaction of the survivors with a device-side count;
utils.bbox_helper.decode_t / clip_t are the same arithmetic as tensor ops)
-> scda_nms_dyn (count read on the device, scan stops after post_nms_top_n survivors) -> gather.

`compute_rpn_proposals` keeps the reference's signature and return type (CPU float tensor
[N, 6] = batch, x1, y1, x2, y2, score); `rpn_proposals_device` is the same computation
returning fixed-capacity device buffers plus counts for the in-graph training path.
"""
import torch

from .._lib import check, load, stream_ptr
from ..extensions._nms.pth_nms import nms_device
from ..utils import anchor_helper


import os as _os
_FORCE_TOPK = _os.environ.get('SCDA_RPN_TOPK', '0') == '1'      # debugging switch


def _image_hw(image_info, b):
    if torch.is_tensor(image_info):
        if image_info.is_cuda:
            image_info = image_info.cpu()      # [B,3] shape info; the driver keeps it on the host
        image_info = image_info.numpy()
    return float(image_info[b][0]), float(image_info[b][1])


def rpn_proposals_device(conv_cls, conv_loc, cfg, image_info, fg_scores=None):
    """Returns a list (one per image) of (boxes5 float32 [cap, 5] = x1,y1,x2,y2,score sorted
    by descending score with the NMS survivors first, n_keep int64 0-dim tensor).
    fg_scores: [B, K*A] foreground probabilities already in anchor order (loss_ops.rpn_fg_scores on the
    raw class map) — then conv_cls is not read."""
    assert conv_loc.is_cuda and (fg_scores is not None or conv_cls.is_cuda)
    B, A4, fh, fw = conv_loc.shape
    A = A4 // 4
    assert A * 4 == A4
    KA = fh * fw * A
    dev = conv_loc.device
    anchors = anchor_helper.anchors_device(fh, fw, cfg['anchor_ratios'], cfg['anchor_scales'],
                                           cfg['anchor_stride'], dev)
    cls_view = conv_cls.permute(0, 2, 3, 1).reshape(B, KA, -1) if fg_scores is None else None
    loc_view = conv_loc.permute(0, 2, 3, 1).reshape(B, KA, 4)
    pre, post = cfg['pre_nms_top_n'], cfg['post_nms_top_n']
    out = []
    for b in range(B):
        scores = cls_view[b, :, -1].contiguous() if fg_scores is None else fg_scores[b]
        h, w = _image_hw(image_info, b)
        deltas = loc_view[b].float().contiguous()
        n = KA if (pre <= 0 or pre > KA) else pre
        packed = torch.empty(n, 5, dtype=torch.float32, device=dev)
        count = torch.empty(1, dtype=torch.int32, device=dev)
        if KA <= 51200 and not _FORCE_TOPK:
            # selection, descending order, decode, clip, min-size filter, compaction: two launches
            # (csrc/proposal_ops.cu)
            lib = load()
            scores = scores.float().contiguous()
            wsb = lib.scda_rpn_proposal_rows_workspace_bytes(KA, pre)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                check(lib.scda_rpn_proposal_rows(KA, pre, scores.data_ptr(), anchors.data_ptr(), deltas.data_ptr(),
                                                 float(h), float(w), float(cfg['roi_min_size']), packed.data_ptr(),
                                                 count.data_ptr(), ws.data_ptr(), wsb, stream_ptr(dev)),
                      "scda_rpn_proposal_rows")
        else:
            # more anchors than the selection kernel's shared memory holds: library top-k, then the decode kernel
            from .. import gan_ops
            gan_ops.note_library_call("rpn top-k", ": more than 51200 anchors per image, torch.topk / torch.sort")
            if pre <= 0 or pre > KA:
                top, order = torch.sort(scores, descending=True)
            else:
                top, order = torch.topk(scores, pre, sorted=True)
            with torch.cuda.device(dev):
                check(load().scda_rpn_decode_pack(n, anchors.data_ptr(), deltas.data_ptr(), order.data_ptr(),
                                                  top.float().contiguous().data_ptr(), float(h), float(w),
                                                  float(cfg['roi_min_size']), packed.data_ptr(), count.data_ptr(),
                                                  stream_ptr(dev)), "scda_rpn_decode_pack")
        keep, n_keep = nms_device(packed, cfg['nms_iou_thresh'], max_keep=max(post, 0),
                                  n_dev=count)
        cap = min(post, n) if post > 0 else n
        sel = packed[keep[:cap].clamp(min=0, max=n - 1)]
        valid = torch.arange(cap, device=dev) < n_keep
        sel = torch.where(valid.unsqueeze(1), sel, torch.zeros_like(sel))
        out.append((sel, n_keep.reshape(())))
    return out


def compute_rpn_proposals(conv_cls, conv_loc, cfg, image_info):
    '''
    :argument
        cfg: configs
        conv_cls: FloatTensor, [batch, num_anchors * x, h, w], conv output of classification
        conv_loc: FloatTensor, [batch, num_anchors * 4, h, w], conv output of localization
        image_info: FloatTensor, [batch, 3], image size
    :returns
        proposals: FloatTensor (CPU), [N, 6]: batch_ix, x1, y1, x2, y2, score
    '''
    res = rpn_proposals_device(conv_cls, conv_loc, cfg, image_info)
    rows = []
    for b, (sel, n_keep) in enumerate(res):
        k = int(n_keep.item())
        p = sel[:k]
        rows.append(torch.cat([torch.full((k, 1), float(b), device=p.device), p], dim=1))
    batch_proposals = torch.cat(rows, dim=0).float().cpu()
    return batch_proposals
