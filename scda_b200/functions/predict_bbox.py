"""Final detections at test time (mirrors functions/predict_bbox.py:13-66 of the reference).

Reference: D2H of rois / class probabilities / deltas, then per class: numpy decode, clip,
argsort, H2D, GPU NMS mask, D2H, host scan — eight NMS round trips per image — and a numpy
top-n.  Here decode, per-class sort, NMS (scda_nms, device scan) and the top-n all stay on
the device; one synchronisation at the end sizes the returned tensor.
"""
import torch

from ..extensions._nms.pth_nms import nms_device
from ..utils.bbox_helper import clip_t, decode_t


def compute_predicted_bboxes(rois, pred_cls, pred_loc, image_info, cfg):
    '''
    :param rois: [N, k] k>=5, batch_ix, x1, y1, x2, y2
    :param pred_cls: [N, num_classes]   (softmax probabilities)
    :param pred_loc: [N, num_classes * 4]
    :param image_info: [B, 3]
    :return: bboxes: CUDA float [M, 7]: batch_ix, x1, y1, x2, y2, score, cls
    '''
    dev = pred_cls.device
    assert dev.type == "cuda"
    rois = rois.to(dev).float()
    if torch.is_tensor(image_info):
        image_info = image_info.cpu().numpy()
    N, num_classes = pred_cls.shape[0:2]
    B = len(image_info) if N == 0 else int(rois[:, 0].max().item()) + 1
    means = torch.tensor(cfg['bbox_normalize_means'], dtype=torch.float64, device=dev)
    stds = torch.tensor(cfg['bbox_normalize_stds'], dtype=torch.float64, device=dev)
    per_image = []
    for b in range(B):
        in_b = rois[:, 0] == b
        rb = rois[in_b][:, 1:5]
        nb = rb.shape[0]
        if nb == 0:
            continue
        h, w = float(image_info[b][0]), float(image_info[b][1])
        rows, valid = [], []
        for cls in range(1, num_classes):
            scores = pred_cls[in_b][:, cls].float()
            deltas = pred_loc[in_b][:, cls * 4:cls * 4 + 4].float()
            if cfg['bbox_normalize_stats_precomputed']:
                deltas64 = deltas.double() * stds + means
                # np.exp of the float64 de-normalised deltas is float64 here (:31-33)
                bw, bh = (rb[:, 2] - rb[:, 0]).double(), (rb[:, 3] - rb[:, 1]).double()
                cx, cy = ((rb[:, 0] + rb[:, 2]) / 2.).double(), ((rb[:, 1] + rb[:, 3]) / 2.).double()
                ncx, ncy = deltas64[:, 0] * bw + cx, deltas64[:, 1] * bh + cy
                nw, nh = torch.exp(deltas64[:, 2]) * bw, torch.exp(deltas64[:, 3]) * bh
                boxes = torch.stack([ncx - nw / 2., ncy - nh / 2., ncx + nw / 2., ncy + nh / 2.], 1)
            else:
                boxes = decode_t(rb, deltas).double()
            boxes = clip_t(boxes, h, w)
            ok = scores > cfg['score_thresh'] if cfg['score_thresh'] > 0 else torch.ones_like(scores, dtype=torch.bool)
            key = torch.where(ok, scores, torch.full_like(scores, -1.0))
            s_sorted, order = torch.sort(key, descending=True)
            n_live = ok.sum().to(torch.int32).reshape(1)
            dets = torch.cat([boxes[order].float(), s_sorted.unsqueeze(1)], 1).contiguous()
            keep, n_keep = nms_device(dets, cfg['nms_iou_thresh'], n_dev=n_live)
            kept = dets[keep[:nb].clamp(min=0, max=nb - 1)]
            rows.append(torch.cat([torch.full((nb, 1), float(b), device=dev), kept,
                                   torch.full((nb, 1), float(cls), device=dev)], 1))
            valid.append(torch.arange(nb, device=dev) < n_keep)
        rows = torch.cat(rows, 0)
        valid = torch.cat(valid, 0)
        score_key = torch.where(valid, rows[:, 5], torch.full_like(rows[:, 5], -1.0))
        n_valid = valid.sum()
        if cfg['top_n'] > 0:
            k = min(cfg['top_n'], rows.shape[0])
            _, top = torch.topk(score_key, k, sorted=True)
            rows = rows[top]
            n_valid = torch.clamp(n_valid, max=k)
        else:
            rows = rows[torch.sort(score_key, descending=True)[1]]
        per_image.append((rows, n_valid))
    out = [r[:int(n.item())] for r, n in per_image]
    if not out:
        return torch.zeros(0, 7, device=dev)
    return torch.cat(out, 0).float()
