"""Final detections at test time (mirrors functions/predict_bbox.py:13-66 of the reference).

One image per call with a top-n (the evaluation loop's case): three launches — csrc/predict_ops.cu
(threshold + sort + float64 decode per class; ranking of the survivors) around ONE scda_nms_groups — and one
synchronisation for the number of rows (`_predict_one_image`).  Batches of several images / no top-n: the
tensor-op form below.

Reference: D2H of rois / class probabilities / deltas, then per class and per image: numpy decode,
clip, argsort, H2D, GPU NMS mask, D2H, host scan — eight NMS round trips per image — and a numpy
top-n.  Here all classes of an image are decoded at once, sorted with ONE batched sort and suppressed by
ONE launch of scda_nms_groups (csrc/nms.cu: one CTA per class, mask and scan in shared memory); the
top-n stays on the device and a single synchronisation at the end sizes the returned tensor.

Row order of the result = the reference's: with cfg['top_n'] > 0 each image's rows by descending score;
otherwise class-major (class outer, image inner), each (class, image) group by descending score.
"""
import torch

from .._lib import check, load, stream_ptr
from ..utils.bbox_helper import clip_t


# tests: True sends single-image calls through the tensor-op form too (the two forms are compared)
FORCE_TENSOR_PATH = False


def _decode_all_classes(rb, pred_loc, num_classes, cfg, dev):
    """rb [n, 4] fp32, pred_loc [n, 4 * num_classes] -> float64 boxes [num_classes - 1, n, 4] (classes 1..)
    with the reference's arithmetic (float64 de-normalised deltas, centre / size without +1)"""
    n = rb.shape[0]
    deltas = pred_loc.float().view(n, num_classes, 4)[:, 1:].permute(1, 0, 2).double()       # [Cc, n, 4]
    if cfg['bbox_normalize_stats_precomputed']:
        stds = torch.tensor(cfg['bbox_normalize_stds'], dtype=torch.float64, device=dev)
        means = torch.tensor(cfg['bbox_normalize_means'], dtype=torch.float64, device=dev)
        deltas = deltas * stds + means
    else:
        deltas = deltas.float().double()
    bw, bh = (rb[:, 2] - rb[:, 0]).double(), (rb[:, 3] - rb[:, 1]).double()
    cx, cy = ((rb[:, 0] + rb[:, 2]) / 2.).double(), ((rb[:, 1] + rb[:, 3]) / 2.).double()
    ncx, ncy = deltas[..., 0] * bw + cx, deltas[..., 1] * bh + cy
    nw, nh = torch.exp(deltas[..., 2]) * bw, torch.exp(deltas[..., 3]) * bh
    return torch.stack([ncx - nw / 2., ncy - nh / 2., ncx + nw / 2., ncy + nh / 2.], 2)


def nms_groups(dets, n_live, thresh):
    """dets fp32 [G, n, 5] (each group sorted by descending score), n_live int32 [G] -> (keep int64 [G, n]
    ascending survivor indices, counts int64 [G]): ONE launch for all groups"""
    G, n, _ = dets.shape
    keep = torch.zeros(G, n, dtype=torch.int64, device=dets.device)
    num = torch.zeros(G, dtype=torch.int64, device=dets.device)
    with torch.cuda.device(dets.device):
        check(load().scda_nms_groups(G, n, n_live.data_ptr(), dets.data_ptr(), float(thresh), keep.data_ptr(),
                                     num.data_ptr(), stream_ptr(dets.device)), "scda_nms_groups")
    return keep, num


def _predict_one_image(rois, pred_cls, pred_loc, info_row, cfg):
    """One image, cfg['top_n'] > 0: three launches (csrc/predict_ops.cu around scda_nms_groups) and one
    synchronisation for the number of rows.  -> CUDA float [M, 7]"""
    import ctypes as C
    dev = pred_cls.device
    n, num_classes = pred_cls.shape[0:2]
    Cc, top_n = num_classes - 1, int(cfg['top_n'])
    rois = rois.contiguous()
    cls = pred_cls.reshape(n, num_classes).float().contiguous()
    loc = pred_loc.reshape(n, num_classes * 4).float().contiguous()
    dets = torch.empty(Cc, n, 5, dtype=torch.float32, device=dev)
    n_live = torch.empty(Cc, dtype=torch.int32, device=dev)
    norm = bool(cfg['bbox_normalize_stats_precomputed'])
    stds = (C.c_double * 4)(*[float(v) for v in cfg['bbox_normalize_stds']]) if norm else None
    means = (C.c_double * 4)(*[float(v) for v in cfg['bbox_normalize_means']]) if norm else None
    lib, st = load(), stream_ptr(dev)
    with torch.cuda.device(dev):
        check(lib.scda_predict_prepare(n, num_classes, rois.data_ptr(), rois.stride(0), cls.data_ptr(), loc.data_ptr(),
                                       1 if norm else 0, stds, means, float(info_row[0]), float(info_row[1]),
                                       float(cfg['score_thresh']), dets.data_ptr(), n_live.data_ptr(), st),
              "scda_predict_prepare")
    keep, n_keep = nms_groups(dets, n_live, cfg['nms_iou_thresh'])
    rows = torch.empty(top_n, 7, dtype=torch.float32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.scda_predict_topn(Cc, n, dets.data_ptr(), keep.data_ptr(), n_keep.data_ptr(), 0.0, top_n,
                                    rows.data_ptr(), count.data_ptr(), st), "scda_predict_topn")
    return rows[:int(count.item())]


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
    Cc = num_classes - 1
    if (len(image_info) == 1 and 0 < N <= 1024 and cfg['top_n'] > 0 and 2 <= num_classes <= 65
            and rois.shape[1] >= 5 and pred_cls.dim() <= 4 and not FORCE_TENSOR_PATH):
        # one image per call (the reference's evaluation loop runs batch size 1 per GPU): every RoI is image 0's
        return _predict_one_image(rois, pred_cls, pred_loc, image_info[0], cfg)
    B = len(image_info) if N == 0 else int(rois[:, 0].max().item()) + 1
    per_image = []          # (rows [Cc, nb, 7], valid [Cc, nb])
    for b in range(B):
        in_b = rois[:, 0] == b
        rb = rois[in_b][:, 1:5]
        nb = rb.shape[0]
        if nb == 0:
            continue
        if nb > 1024:
            raise ValueError("compute_predicted_bboxes: more than 1024 RoIs per image (%d)" % nb)
        h, w = float(image_info[b][0]), float(image_info[b][1])
        boxes = clip_t(_decode_all_classes(rb, pred_loc[in_b], num_classes, cfg, dev).view(-1, 4), h, w).view(Cc, nb, 4)
        scores = pred_cls[in_b][:, 1:].float().t().contiguous()                                  # [Cc, nb]
        ok = scores > cfg['score_thresh'] if cfg['score_thresh'] > 0 else torch.ones_like(scores, dtype=torch.bool)
        key = torch.where(ok, scores, torch.full_like(scores, -1.0))
        s_sorted, order = torch.sort(key, dim=1, descending=True)                                # one batched sort
        dets = torch.cat([torch.gather(boxes.float(), 1, order.unsqueeze(2).expand(-1, -1, 4)),
                          s_sorted.unsqueeze(2)], 2).contiguous()                                # [Cc, nb, 5]
        n_live = ok.sum(1).to(torch.int32).contiguous()
        keep, n_keep = nms_groups(dets, n_live, cfg['nms_iou_thresh'])
        kept = torch.gather(dets, 1, keep.clamp(min=0, max=nb - 1).unsqueeze(2).expand(-1, -1, 5))
        cls_col = torch.arange(1, num_classes, device=dev, dtype=torch.float32).view(Cc, 1, 1).expand(-1, nb, -1)
        rows = torch.cat([torch.full((Cc, nb, 1), float(b), device=dev), kept, cls_col], 2)
        valid = torch.arange(nb, device=dev).unsqueeze(0) < n_keep.unsqueeze(1)
        per_image.append((rows, valid))
    if not per_image:
        return torch.zeros(0, 7, device=dev)
    if cfg['top_n'] > 0:
        out = []
        for rows, valid in per_image:
            rows, valid = rows.reshape(-1, 7), valid.reshape(-1)
            score_key = torch.where(valid, rows[:, 5], torch.full_like(rows[:, 5], -1.0))
            k = min(cfg['top_n'], rows.shape[0])
            _, top = torch.topk(score_key, k, sorted=True)
            n_valid = torch.clamp(valid.sum(), max=k)
            out.append((rows[top], n_valid))
        parts = [r[:int(n.item())] for r, n in out]
    else:
        # the reference's order without a top-n: class outer, image inner (:27-52)
        parts = []
        masks = [v.cpu() for _, v in per_image]          # one synchronisation for the sizes
        for c in range(Cc):
            for (rows, _), v in zip(per_image, masks):
                parts.append(rows[c][:int(v[c].sum())])
    return torch.cat(parts, 0).float()
