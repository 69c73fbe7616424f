"""numpy restatement of the reference's HOST-side plumbing on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/scda_oracle.c).  The reference runs all of
this in numpy on the host between GPU ops; each function cites the file:line it
follows.  Random draws (`np.random.choice`, unseeded in the reference) go
through an injectable `choice(n, size, replace)` so that parity tests can feed
the device pipeline and this restatement the same decisions.

Pinned against the reference itself: the IoU inside goes through the reference's
real cython_bbox arithmetic (oracle.bbox_overlaps == oracle/_ref cython_bbox,
tests/test_oracle_cpu.py); everything else here is plain numpy that mirrors the
reference's own numpy line by line — there is no other implementation to pin to.
"""
from __future__ import annotations

import numpy as np

from . import bbox_overlaps, nms as _nms


def _pick(choice, tag, n, size, replace):
    """One random draw.  A plain `choice` has np.random.choice's signature; a test adapter
    that sets `wants_tag` also receives which draw this is (the reference's draws are
    conditional, so position in the call sequence does not identify them)."""
    if getattr(choice, "wants_tag", False):
        return choice(n, size=size, replace=replace, tag=tag)
    return choice(n, size=size, replace=replace)


# ------------------------------------------------------------------ anchors
def anchors_over_grid(ratios, scales, stride):
    """utils/anchor_helper.py:4-11,37-96.  NOTE the reference ignores `ratios`
    (get_anchors_over_grid returns before using it, :9-11) and always enumerates
    (0.5, 1, 2); ratio-major, scale-minor order."""
    del ratios
    aspect = np.array((0.5, 1, 2), dtype=float)
    sc = np.array(np.array(scales) * stride, dtype=float) / stride
    base = np.array([1, 1, stride, stride], dtype=float) - 1
    w = base[2] - base[0] + 1
    h = base[3] - base[1] + 1
    xc = base[0] + 0.5 * (w - 1)
    yc = base[1] + 0.5 * (h - 1)
    ws = np.round(np.sqrt(w * h / aspect))
    hs = np.round(ws * aspect)
    out = []
    for wi, hi in zip(ws, hs):
        for s in sc:
            ww, hh = wi * s, hi * s
            out.append([xc - 0.5 * (ww - 1), yc - 0.5 * (hh - 1), xc + 0.5 * (ww - 1), yc + 0.5 * (hh - 1)])
    return np.array(out)


def anchors_over_plane(fh, fw, ratios, scales, stride):
    """utils/anchor_helper.py:21-35: [K*A, 4] float64, cell-major (y, x), anchor-minor."""
    grid = anchors_over_grid(ratios, scales, stride)
    sx, sy = np.meshgrid(np.arange(0, fw) * stride, np.arange(0, fh) * stride)
    shifts = np.vstack((sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel())).transpose()
    A, K = grid.shape[0], shifts.shape[0]
    return (grid.reshape((1, A, 4)) + shifts.reshape((1, K, 4)).transpose((1, 0, 2))).reshape((K * A, 4))


# ---------------------------------------------------------------- box codec
def corner_to_center(b):
    """utils/bbox_helper.py:50-58 (no +1 widths)."""
    return np.vstack([(b[:, 0] + b[:, 2]) / 2., (b[:, 1] + b[:, 3]) / 2.,
                      b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]]).transpose()


def center_to_corner(b):
    """utils/bbox_helper.py:37-48."""
    return np.vstack([b[:, 0] - b[:, 2] / 2., b[:, 1] - b[:, 3] / 2.,
                      b[:, 0] + b[:, 2] / 2., b[:, 1] + b[:, 3] / 2.]).transpose()


def compute_loc_targets(raw, gt):
    """utils/bbox_helper.py:60-76."""
    bb, g = corner_to_center(raw), corner_to_center(gt)
    return np.vstack([(g[:, 0] - bb[:, 0]) / bb[:, 2], (g[:, 1] - bb[:, 1]) / bb[:, 3],
                      np.log(g[:, 2] / bb[:, 2]), np.log(g[:, 3] / bb[:, 3])]).transpose()


def compute_loc_bboxes(raw, deltas):
    """utils/bbox_helper.py:79-96.  raw float64 anchors x float32 deltas: the
    products promote to float64 but np.exp(float32) stays float32."""
    bb = corner_to_center(raw)
    cx = deltas[:, 0] * bb[:, 2] + bb[:, 0]
    cy = deltas[:, 1] * bb[:, 3] + bb[:, 1]
    w = np.exp(deltas[:, 2]) * bb[:, 2]
    h = np.exp(deltas[:, 3]) * bb[:, 3]
    return center_to_corner(np.vstack([cx, cy, w, h]).transpose())


def clip_bbox(b, img_size):
    """utils/bbox_helper.py:105-111 (in place)."""
    h, w = img_size[:2]
    b[:, 0] = np.clip(b[:, 0], 0, w - 1)
    b[:, 1] = np.clip(b[:, 1], 0, h - 1)
    b[:, 2] = np.clip(b[:, 2], 0, w - 1)
    b[:, 3] = np.clip(b[:, 3], 0, h - 1)
    return b


def bbox_iou_overlaps(b1, b2):
    """utils/bbox_helper.py:8-9 -> cython_bbox.bbox_overlaps on float32 casts."""
    return bbox_overlaps(np.ascontiguousarray(b1[:, :4].astype(np.float32)),
                         np.ascontiguousarray(b2[:, :4].astype(np.float32)))


# ------------------------------------------------------------ rpn proposals
def compute_rpn_proposals(conv_cls, conv_loc, cfg, image_info):
    """functions/rpn_proposal.py:17-74.  conv_cls [B, A*2, H, W] (softmax probabilities),
    conv_loc [B, A*4, H, W], numpy.  Returns float32 [N, 6] (b, x1, y1, x2, y2, score)."""
    B, A4, fh, fw = conv_loc.shape
    anchors = anchors_over_plane(fh, fw, cfg['anchor_ratios'], cfg['anchor_scales'], cfg['anchor_stride'])
    A, K = A4 // 4, fh * fw
    cls_view = np.ascontiguousarray(conv_cls.transpose(0, 2, 3, 1)).reshape(B, K * A, -1)
    loc_view = np.ascontiguousarray(conv_loc.transpose(0, 2, 3, 1)).reshape(B, K * A, 4)
    out = []
    pre = cfg['pre_nms_top_n']
    for b in range(B):
        scores = cls_view[b, :, -1]
        if pre <= 0 or pre > scores.shape[0]:
            order = scores.argsort()[::-1]
        else:
            inds = np.argpartition(-scores, pre)[:pre]
            order = inds[np.argsort(-scores[inds])]
        boxes = compute_loc_bboxes(anchors[order, :], loc_view[b, order, :])
        boxes = clip_bbox(boxes, image_info[b])
        props = np.hstack([boxes, scores[order][:, np.newaxis]])
        props = props[(props[:, 2] - props[:, 0] + 1 >= cfg['roi_min_size'])
                      & (props[:, 3] - props[:, 1] + 1 >= cfg['roi_min_size'])]
        keep = _nms(props.astype(np.float32), cfg['nms_iou_thresh'])
        if cfg['post_nms_top_n'] > 0:
            keep = keep[:cfg['post_nms_top_n']]
        props = props[keep]
        out.append(np.hstack([np.full((len(keep), 1), b, dtype=props.dtype), props]))
    return np.vstack(out).astype(np.float32)


# ------------------------------------------------------------ anchor targets
def compute_anchor_targets(feature_size, cfg, gts, image_info, choice=None):
    """functions/anchor_target.py:16-116 (ignore_regions=None, as the driver passes,
    tools/faster_rcnn_train_val.py:520).  Returns numpy cls_targets [B,A,H,W] int64,
    loc_targets / loc_masks [B,4A,H,W] float32, normalizer."""
    choice = choice or np.random.choice
    B, A4, fh, fw = feature_size
    A, K = A4 // 4, fh * fw
    anchors = anchors_over_plane(fh, fw, cfg['anchor_ratios'], cfg['anchor_scales'], cfg['anchor_stride'])
    overlaps = np.stack([bbox_iou_overlaps(anchors, gts[i]) for i in range(B)], axis=0)
    argmax = overlaps.argmax(axis=2)
    mx = overlaps.max(axis=2)
    gt_max = overlaps.max(axis=1)
    gt_max[gt_max < 0.1] = -1
    gb, gka, gg = np.where(overlaps == gt_max[:, np.newaxis, :])
    argmax[gb, gka] = gg
    labels = np.full([B, K * A], -1, dtype=np.int64)
    labels[mx < cfg['negative_iou_thresh']] = 0
    labels[gb, gka] = 1
    labels[mx > cfg['positive_iou_thresh']] = 1
    n_pos_want = int(cfg['positive_percent'] * cfg['rpn_batch_size'] * B)
    pb, pka = np.where(labels > 0)
    n_pos = len(pb)
    if n_pos > n_pos_want:
        rm = _pick(choice, 'pos', n_pos, n_pos - n_pos_want, False)
        labels[pb[rm], pka[rm]] = -1
        n_pos = n_pos_want
    n_neg_want = cfg['rpn_batch_size'] * B - n_pos
    nb, nka = np.where(labels == 0)
    if len(nb) > n_neg_want:
        rm = _pick(choice, 'neg', len(nb), len(nb) - n_neg_want, False)
        labels[nb[rm], nka[rm]] = -1
    pb, pka = np.where(labels > 0)
    tgt = compute_loc_targets(anchors[pka, :], gts[pb, argmax[pb, pka]])
    loc_t = np.zeros([B, K * A, 4], dtype=np.float32)
    loc_t[pb, pka, :] = tgt
    loc_m = np.zeros([B, K * A, 4], dtype=np.float32)
    loc_m[pb, pka, :] = 1.
    cls_targets = np.ascontiguousarray(labels.reshape(B, fh, fw, A).transpose(0, 3, 1, 2))
    loc_targets = np.ascontiguousarray(loc_t.reshape(B, fh, fw, A * 4).transpose(0, 3, 1, 2))
    loc_masks = np.ascontiguousarray(loc_m.reshape(B, fh, fw, A * 4).transpose(0, 3, 1, 2))
    return cls_targets, loc_targets, loc_masks, max(1, int((labels >= 0).sum()))


# ---------------------------------------------------------- proposal targets
def compute_proposal_targets(proposals, cfg, gts_all, image_info, choice=None):
    """functions/proposal_target.py:17-177 (ignore_regions=None, use_ohem=False).
    proposals [N, >=5] numpy; returns rois [n,5] f32, labels [n] i64, loc_targets /
    loc_weights [n, num_classes*4] f32."""
    choice = choice or np.random.choice
    B = gts_all.shape[0]
    o_rois, o_lab, o_t, o_w = [], [], [], []
    for b in range(B):
        rois = proposals[proposals[:, 0] == b][:, 1:5]
        gts = gts_all[b]
        gts = gts[(gts[:, 2] > gts[:, 0] + 1) & (gts[:, 3] > gts[:, 1] + 1)]
        if cfg['append_gts']:
            rois = np.vstack([rois, gts[:, :4]])
        rois = clip_bbox(rois, image_info[b])
        if rois.shape[0] == 0 or gts.shape[0] == 0:
            continue
        ov = bbox_iou_overlaps(rois, gts)
        amax, mx = ov.argmax(axis=1), ov.max(axis=1)
        pos = np.where(mx > cfg['positive_iou_thresh'])[0]
        pos_g = amax[pos]
        neg = np.where((mx < cfg['negative_iou_thresh_hi']) & (mx >= cfg['negative_iou_thresh_lo']))[0]
        # np.array(list(set(neg) - set(pos))) in the reference (:90): pos and neg are disjoint
        # when positive_iou_thresh >= negative_iou_thresh_hi, and the set round-trip then keeps
        # the values; small non-negative ints hash to themselves, so iteration order is
        # ascending for the sizes seen here.  Written as a sorted difference.
        neg = np.array(sorted(set(neg.tolist()) - set(pos.tolist())), dtype=np.int64)
        n_pos = len(pos)
        bs = cfg['batch_size']
        want_pos = int(cfg['positive_percent'] * bs)
        if want_pos < n_pos:
            k = _pick(choice, 'pos', n_pos, want_pos, False)
            pos, pos_g, n_pos = pos[k], pos_g[k], want_pos
        want_neg = bs - n_pos
        if want_neg < len(neg):
            k = _pick(choice, 'neg', len(neg), want_neg, False)
            neg = neg[k]
        pos_rois, pos_gts, neg_rois = rois[list(pos)], gts[list(pos_g)], rois[list(neg)]
        sampled = np.vstack([pos_rois, neg_rois])
        npos, nneg = pos_rois.shape[0], neg_rois.shape[0]
        pos_labels = pos_gts[:, 4].astype(np.int32)
        labels = np.concatenate([pos_labels, np.zeros(nneg)]).astype(np.int32)
        loc_t = np.zeros([npos + nneg, cfg['num_classes'], 4])
        loc_w = np.zeros([npos + nneg, cfg['num_classes'], 4])
        t = compute_loc_targets(pos_rois, pos_gts)
        if cfg['bbox_normalize_stats_precomputed']:
            t = (t - np.array(cfg['bbox_normalize_means'])[np.newaxis, :]) / np.array(cfg['bbox_normalize_stds'])[np.newaxis, :]
        loc_t[range(npos), pos_labels, :] = t
        loc_w[range(npos), pos_labels, :] = 1
        loc_t = loc_t.reshape([npos + nneg, -1])
        loc_w = loc_w.reshape([npos + nneg, -1])
        sampled = np.hstack([np.full((sampled.shape[0], 1), b, dtype=sampled.dtype), sampled])
        if sampled.shape[0] < bs:
            rep = _pick(choice, 'pad', sampled.shape[0], bs - sampled.shape[0], True)
            sampled = np.vstack([sampled, sampled[rep]])
            labels = np.concatenate([labels, labels[rep]])
            loc_t = np.vstack([loc_t, loc_t[rep]])
            loc_w = np.vstack([loc_w, loc_w[rep]])
        o_rois.append(sampled); o_lab.append(labels); o_t.append(loc_t); o_w.append(loc_w)
    return (np.vstack(o_rois).astype(np.float32), np.concatenate(o_lab).astype(np.int64),
            np.vstack(o_t).astype(np.float32), np.vstack(o_w).astype(np.float32))


# --------------------------------------------------------- predicted bboxes
def compute_predicted_bboxes(rois, pred_cls, pred_loc, image_info, cfg):
    """functions/predict_bbox.py:13-66: per class decode + clip + sort + NMS + top-n.
    Returns float32 [M, 7] (b, x1, y1, x2, y2, score, cls)."""
    N, num_classes = pred_cls.shape[0:2]
    B = int(max(rois[:, 0].astype(np.int32)) + 1)
    res = []
    for cls in range(1, num_classes):
        scores = pred_cls[:, cls]
        deltas = pred_loc[:, cls * 4:cls * 4 + 4]
        if cfg['bbox_normalize_stats_precomputed']:
            deltas = deltas * np.array(cfg['bbox_normalize_stds'])[np.newaxis, :] \
                + np.array(cfg['bbox_normalize_means'])[np.newaxis, :]
        bboxes = np.hstack([compute_loc_bboxes(rois[:, 1:5], deltas), scores[:, np.newaxis]])
        for b in range(B):
            ix = np.where(rois[:, 0] == b)[0]
            ps, pb = scores[ix], bboxes[ix]
            pb[:, :4] = clip_bbox(pb[:, :4], image_info[b])
            if cfg['score_thresh'] > 0:
                k = np.where(ps > cfg['score_thresh'])[0]
                ps, pb = ps[k], pb[k]
            if ps.size == 0:
                continue
            pb = pb[ps.argsort()[::-1], :]
            keep = _nms(pb.astype(np.float32), cfg['nms_iou_thresh'])
            post = pb[keep]
            res.append(np.hstack([np.full((len(keep), 1), b), post, np.full((len(keep), 1), cls)]))
    res = np.vstack(res)
    if cfg['top_n'] > 0:
        tops = []
        for b in range(B):
            bb = res[res[:, 0] == b]
            tops.append(bb[bb[:, -2].argsort()[::-1][:cfg['top_n']]])
        res = np.vstack(tops)
    return res.astype(np.float32)


# ------------------------------------------------------------------- losses
def smooth_l1_loss_with_sigma(pred, targets, sigma=3.0):
    """models/faster_rcnn/faster_rcnn_adver_expansion_reweight_cluster.py:238-246 (sum)."""
    s2 = sigma ** 2
    d = pred.astype(np.float32) - targets.astype(np.float32)
    a = np.abs(d)
    sign = (a < 1. / s2).astype(np.float32)
    return float(np.sum(d * d * s2 / 2. * sign + (a - 0.5 / s2) * (1. - sign), dtype=np.float64))


def cross_entropy(logits, targets, ignore_index=-1):
    """F.cross_entropy(..., ignore_index) as used at :51-52 and :63: mean over kept rows."""
    x = logits.astype(np.float64)
    x = x - x.max(1, keepdims=True)
    lse = np.log(np.exp(x).sum(1))
    keep = targets != ignore_index
    nll = lse[keep] - x[keep, targets[keep]]
    return float(nll.mean()) if keep.any() else float('nan')


def accuracy_top1(logits, targets, ignore_index=-1):
    """:249-267 with topk=(1,): percentage of kept rows whose argmax is the target."""
    keep = targets != ignore_index
    return float((logits[keep].argmax(1) == targets[keep]).mean() * 100.0)


# -------------------------------------------------------------- region crops
def get_corner_from_center(center, recon_size, new_w, new_h):
    """tools/faster_rcnn_train_val.py:411-438."""
    half = recon_size // 2
    out = []
    for cx, cy in center:
        x1 = max(int(cx) - half, 0)
        y1 = max(int(cy) - half, 0)
        if x1 == 0:
            x2 = recon_size
        else:
            x2 = min(int(cx) + half, new_w)
            if x2 == new_w:
                x1 = new_w - recon_size
        if y1 == 0:
            y2 = recon_size
        else:
            y2 = min(int(cy) + half, new_h)
            if y2 == new_h:
                y1 = new_h - recon_size
        out.append([x1, y1, x2, y2])
    return out


def compute_cluster_targets(proposals, features, n_cluster=4, threshold=128, choice=None):
    """functions/mask.py:183-237: sklearn KMeans(random_state=0) on the RoI centres, then
    `threshold` feature rows per cluster (first members, or resampled with replacement)."""
    from sklearn.cluster import KMeans
    choice = choice or np.random.choice
    centers = np.vstack([(proposals[:, 3] + proposals[:, 1]) / 2.0,
                         (proposals[:, 4] + proposals[:, 2]) / 2.0]).transpose()
    km = KMeans(n_clusters=n_cluster, random_state=0).fit(centers)
    out = []
    for c in range(n_cluster):
        ix = np.where(km.labels_ == c)[0]
        if ix.shape[0] < threshold:
            ix = ix[choice(ix.shape[0], threshold, replace=True)]
        else:
            ix = ix[:threshold]
        out.append(features[ix])
    return np.stack(out, axis=0).astype(np.float32), km.cluster_centers_, km.labels_
