"""Cityscapes-style detection mAP from result lines and the ground-truth meta file (mirrors
utils/cal_mAP.py:16-174 of the reference: VOC-style AP with +1 box areas, IoU >= 0.5, each ground truth
matched at most once, the mean over classes 1..num_classes-1 — the number README.md:94 quotes, 33.91).
Pure numpy / Python: evaluation plumbing, not on the training hot path."""
from collections import defaultdict

import numpy as np


def parse_gts(gts_list, num_classes):
    """meta file lines -> dict[img_name] = {'bbox': {cls: [[x1,y1,x2,y2], ...]}, 'is_det': {cls: flags}}"""
    index = [i for i, line in enumerate(gts_list) if line.startswith('#')]
    gts = defaultdict(list)
    gts['num'] = np.zeros(num_classes)
    for ix in index:
        pure_name = gts_list[ix + 1].strip().split('/')[-1][0:-4]
        n_box = int(gts_list[ix + 7])
        ent = {'height': gts_list[ix + 3].strip(), 'width': gts_list[ix + 4].strip(), 'bbox_num': n_box,
               'bbox': defaultdict(list), 'is_det': defaultdict(list)}
        for b in gts_list[ix + 8:ix + 8 + n_box]:
            b = b.split()
            label = int(b[0])
            ent['bbox'][label].append([int(b[1]), int(b[2]), int(b[3]), int(b[4])])
            gts['num'][label] += 1
        for l in range(1, num_classes):
            ent['is_det'][l] = np.zeros(len(ent['bbox'][l]))
        gts[pure_name] = ent
    return gts


def parse_res(res_list):
    """result lines `img x1 y1 x2 y2 score cls` -> dict[cls] = [[x1, y1, x2, y2, score, img], ...]"""
    results = defaultdict(list)
    for r in res_list:
        r = r.split()
        if len(r) < 7:
            continue
        results[int(r[6])].append([int(float(r[1])), int(float(r[2])), int(float(r[3])), int(float(r[4])),
                                   float(r[5]), r[0]])
    return results


def calIoU(result, gt_i):
    x1, y1, x2, y2 = result[:4]
    overmax, is_which = -1, -1
    for k, gt in enumerate(gt_i):
        ix1, iy1, ix2, iy2 = max(x1, gt[0]), max(y1, gt[1]), min(x2, gt[2]), min(y2, gt[3])
        if ix1 < ix2 and iy1 < iy2:
            inter = (ix2 - ix1 + 1) * (iy2 - iy1 + 1)
            iou = inter / ((x2 - x1 + 1) * (y2 - y1 + 1) + (gt[2] - gt[0] + 1) * (gt[3] - gt[1] + 1) - inter)
            if iou > overmax:
                overmax, is_which = iou, k
    return overmax, is_which


def cal_mAP(gts, results, num_classes, overlap_thre):
    ap, max_recall = np.zeros(num_classes), np.zeros(num_classes)
    for c in range(1, num_classes):
        res = sorted(results[c], key=lambda xx: xx[4], reverse=True)
        n = len(res)
        tp, fp = np.zeros(n), np.zeros(n)
        sum_gt = gts['num'][c]
        for k, r in enumerate(res):
            ent = gts[r[-1]]
            gts_i = ent['bbox'][int(c)] if isinstance(ent, dict) else []
            overmax, which = calIoU(r, gts_i)
            if overmax >= overlap_thre and ent['is_det'][c][which] == 0:
                tp[k] = 1
                ent['is_det'][c][which] = 1
            else:
                fp[k] = 1
        if n == 0 or sum_gt == 0:
            continue
        tp, fp = np.cumsum(tp), np.cumsum(fp)
        rec, prec = tp / sum_gt, tp / (tp + fp)
        for v in range(n - 2, -1, -1):
            prec[v] = max(prec[v], prec[v + 1])
        ap[c] = rec[0] * prec[0] + float(np.sum((rec[1:] - rec[:-1]) * prec[1:]))
        max_recall[c] = np.max(rec)
    return ap, max_recall


def Cal_MAP1(res_list, gts_list, num_classes):
    num_classes = int(num_classes)
    ap, _ = cal_mAP(parse_gts(gts_list, num_classes), parse_res(res_list), num_classes, 0.5)
    return float(np.mean(ap[1:]))


def Cal_MAP(res_dir, gts_list_path, num_classes):
    """concatenate results.txt.rank* of `res_dir`, evaluate against the meta file; prints and returns mAP"""
    import glob
    import os
    res = []
    for f in sorted(glob.glob(os.path.join(res_dir, 'results.txt.rank*'))):
        with open(f, 'r', encoding='utf-8') as fh:
            res += fh.readlines()
    with open(os.path.join(res_dir, 'results.txt'), 'w', encoding='utf-8') as fh:
        fh.writelines(res)
    with open(gts_list_path, 'r', encoding='utf-8') as fh:
        gts_list = fh.readlines()
    num_classes = int(num_classes)
    ap, max_recall = cal_mAP(parse_gts(gts_list, num_classes), parse_res(res), num_classes, 0.5)
    mAP = float(np.mean(ap[1:]))
    print('--------------------')
    print('mAP: {}   max recall: {}'.format(mAP, float(np.mean(max_recall[1:]))))
    print('--------------------')
    return mAP
