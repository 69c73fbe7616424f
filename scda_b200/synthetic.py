"""Seeded synthetic inputs of the BASELINE.json configurations, shared by
bench.py, __graft_entry__.smoke(), the golden generators and the tests (tests/_inputs.py
re-exports this module).  numpy only, so the same arrays are produced on every box."""
import numpy as np


def rois_uniform(n, seed, img_w=1024, img_h=1024, wh=(16, 512), batch=1):
    """[n, 5] (b, x1, y1, x2, y2) in image coordinates — config 1 of BASELINE.json."""
    r = np.random.RandomState(seed)
    x1 = r.uniform(0, img_w - 16, n)
    y1 = r.uniform(0, img_h - 16, n)
    w = r.uniform(wh[0], wh[1], n)
    h = r.uniform(wh[0], wh[1], n)
    x2 = np.minimum(x1 + w, img_w - 1)
    y2 = np.minimum(y1 + h, img_h - 1)
    b = r.randint(0, batch, n).astype(np.float64)
    return np.stack([b, x1, y1, x2, y2], 1).astype(np.float32)


def features(shape, seed):
    return np.random.RandomState(seed).standard_normal(shape).astype(np.float32)


def nms_boxes(n, seed, img_w=1024, img_h=512):
    """[n, 5] (x1, y1, x2, y2, score) sorted by descending score — config 5."""
    r = np.random.RandomState(seed)
    cx = r.uniform(0, img_w, n)
    cy = r.uniform(0, img_h, n)
    w = np.exp(r.uniform(np.log(8), np.log(512), n))
    h = np.exp(r.uniform(np.log(8), np.log(512), n))
    x1 = np.clip(cx - w / 2, 0, img_w - 1)
    y1 = np.clip(cy - h / 2, 0, img_h - 1)
    x2 = np.clip(cx + w / 2, 0, img_w - 1)
    y2 = np.clip(cy + h / 2, 0, img_h - 1)
    s = np.sort(r.uniform(0, 1, n))[::-1]
    return np.stack([x1, y1, x2, y2, s], 1).astype(np.float32)


def clustered_boxes(n, seed, n_centres=40, jitter=6.0, img_w=1024, img_h=512):
    """Heavily overlapping boxes (many suppressions, long dependency chains)."""
    r = np.random.RandomState(seed)
    c = r.randint(0, n_centres, n)
    cx0 = r.uniform(50, img_w - 50, n_centres)
    cy0 = r.uniform(50, img_h - 50, n_centres)
    w0 = r.uniform(30, 200, n_centres)
    h0 = r.uniform(30, 200, n_centres)
    cx = cx0[c] + r.normal(0, jitter, n)
    cy = cy0[c] + r.normal(0, jitter, n)
    w = w0[c] * np.exp(r.normal(0, 0.1, n))
    h = h0[c] * np.exp(r.normal(0, 0.1, n))
    x1 = np.clip(cx - w / 2, 0, img_w - 1)
    y1 = np.clip(cy - h / 2, 0, img_h - 1)
    x2 = np.clip(cx + w / 2, 0, img_w - 1)
    y2 = np.clip(cy + h / 2, 0, img_h - 1)
    s = np.sort(r.uniform(0, 1, n))[::-1]
    return np.stack([x1, y1, x2, y2, s], 1).astype(np.float32)


def gt_boxes(g, seed, img_w=1024, img_h=512, num_classes=9):
    """[g, 5] (x1, y1, x2, y2, cls) — config 3."""
    r = np.random.RandomState(seed)
    w = r.uniform(20, 300, g)
    h = r.uniform(20, 200, g)
    x1 = r.uniform(0, img_w - w)
    y1 = r.uniform(0, img_h - h)
    cls = r.randint(1, num_classes, g)
    return np.stack([x1, y1, x1 + w, y1 + h, cls], 1).astype(np.float32)


def focal_inputs(m, k, seed, softmax=False):
    r = np.random.RandomState(seed)
    logits = (r.standard_normal((m, k)) * 3).astype(np.float32)
    lo = -1
    hi = k if softmax else k + 1     # softmax labels 0..k-1, sigmoid labels 0..k (d+1 <-> column d)
    targets = r.randint(lo, hi, m).astype(np.int32)
    return logits, targets


def synth_rpn_outputs(seed, fh=32, fw=64, A=15):
    """RPN head outputs as the proposal stage sees them: class map already soft-maxed
    over each anchor's (bg, fg) channel pair, [1, 2A, fh, fw]; deltas [1, 4A, fh, fw]."""
    import torch
    r = np.random.RandomState(seed)
    logits = r.standard_normal((1, 2 * A, fh, fw)).astype(np.float32) * 2
    x = torch.from_numpy(logits).permute(0, 2, 3, 1).contiguous()
    p = torch.softmax(x.view(-1, 2), dim=1).view_as(x).permute(0, 3, 1, 2).contiguous()
    loc = (r.standard_normal((1, 4 * A, fh, fw)) * 0.3).astype(np.float32)
    return p.numpy(), loc


def load_cfg():
    """The reference's config_512.json with `shared` merged into every section
    (tools/faster_rcnn_train_val.py:183-189), written by tests/golden/make_golden_host.py."""
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs",
                                       "config_512_merged.json")))
