#!/usr/bin/env python
"""BASELINE.json configs 2, 3 and 5 on one GPU (configs[3] is bench.py's headline, configs[0] the
parity plumbing of tests/).  One JSON line per config; `python bench.py --config N` calls in here.

  2  vgg16 Faster R-CNN forward-only (eval) on 1x3x512x1024: whole forward and per-stage us
     (backbone, RPN head, proposals = scores + decode + top-k + NMS, RoIPool + RCNN head, predicted boxes)
  3  detector-only train step (forward on the source image, backward, Adam) bs 1, images/s
  5  NMS sweep 1k-100k score-sorted boxes + IoU 30 720 x G, boxes/s and pairs/s; when oracle/_ref holds the
     reference's own kernels (a measurement comparator, never the product path) their time is put beside ours

CUDA events on the launching stream, L2 flushed (512 MB fill) before every timed repetition, median.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

IMG_H, IMG_W, NUM_GT = 512, 1024, 20


def _flush_buf(dev):
    return torch.zeros(128 * 1024 * 1024, device=dev)


def _time(fn, flush, iters=10, warm=3):
    s = torch.cuda.current_stream()
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        fn()
        b.record(s)
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


def _detector(dev, train):
    from scda_b200 import synthetic
    from scda_b200.models.faster_rcnn.vgg_adver_expansion_cluster import vgg16
    cfg = synthetic.load_cfg()
    torch.manual_seed(0)
    model = vgg16(pretrained=False, cfg=cfg['shared']).to(dev)
    model.train() if train else model.eval()
    r = np.random.RandomState(1000)
    image = torch.from_numpy(r.standard_normal((1, 3, IMG_H, IMG_W)).astype(np.float32)).to(dev)
    gts = torch.from_numpy(synthetic.gt_boxes(NUM_GT, 0, img_w=IMG_W, img_h=IMG_H)[None]).to(dev)
    info = torch.tensor([[IMG_H, IMG_W, 0.5]])
    return cfg, model, image, gts, info


def config2(dev):
    """forward-only, per-stage (models/faster_rcnn/…reweight_cluster.py:217-233 of the reference)"""
    import torch.nn.functional as F
    from scda_b200.engine import FlatAdam
    cfg, model, image, gts, info = _detector(dev, train=False)
    FlatAdam(model, 1e-5, tensor_core=True)           # flat storage + bf16 shadows: the tensor-core path
    flush = _flush_buf(dev)
    fns = model._pin_args_to_fn(cfg, None, info, None)
    st = {}

    def backbone():
        st['x'] = model.feature_extractor(image)

    def rpn_head():
        st['cls'], st['loc'] = model.rpn(st['x'])

    def proposals():
        p = fns['rpn_proposal_fn'](model._rpn_scores(st['cls']).data, st['loc'].data)
        st['props'] = p[:, :5].cuda().contiguous()

    def rcnn():
        st['fea'], c, st['ploc'] = model.rcnn(st['x'], st['props'])
        st['pcls'] = F.softmax(c, dim=1)

    def predict():
        st['boxes'] = fns['predict_bbox_fn'](st['props'], st['pcls'], st['ploc'])

    def whole():
        x = {"cfg": cfg, "image": image, "image_info": info, "ground_truth_bboxes": None, "ignore_regions": None}
        st['out'] = model(x)
    stages = {}
    with torch.no_grad():
        for name, fn in (("backbone", backbone), ("rpn_head", rpn_head), ("proposals_nms", proposals),
                         ("roipool_rcnn_head", rcnn), ("predict_bboxes_nms", predict)):
            stages[name + "_us"] = round(_time(fn, flush), 1)
        total = _time(whole, flush)
    return {"config": 2, "workload": "vgg16 Faster R-CNN forward-only (eval), 1x3x512x1024, 1xB200",
            "forward_us": round(total, 1), "images_per_s": 1e6 / total, "stages": stages,
            "proposals": int(st['props'].shape[0]), "detections": int(st['boxes'].shape[0]),
            "dtype": "bf16", "timing": "CUDA events, eager launches (the eval path returns host-sized outputs), "
                                       "L2 flushed before each repetition, median of 10"}


def config3(dev, steps=30):
    """detector-only train step: forward on one labelled image, backward, Adam (the reference's plain
    Faster R-CNN step = its SCDA step without the target image and the reconstruction networks)"""
    from scda_b200.engine import FlatAdam
    from scda_b200.functions.rpn_proposal import rpn_proposals_device
    from scda_b200.loss_ops import rpn_fg_scores
    cfg, model, image, gts, info = _detector(dev, train=True)
    opt = FlatAdam(model, 1.25e-5, weight_decay=1e-4, tensor_core=True)
    pcfg = cfg['train_rpn_proposal_cfg']
    out = {}

    def step():
        fns = model._pin_args_to_fn(cfg, gts, info, None)
        x = model.feature_extractor(image)
        cls, loc = model.rpn(x)
        props = rpn_proposals_device(None, loc.data, pcfg, info, fg_scores=rpn_fg_scores(cls))
        rois, ct, lt, lw = model._train_rois(cfg, props, gts, info)
        _, pc, pl = model.rcnn(x, rois)
        l1, l2, _ = model._add_rpn_loss(fns['anchor_target_fn'], cls, loc)
        l3, l4, _ = model._add_rcnn_loss(pc, pl, ct, lt, lw)
        loss = l1 + l2 + l3 + l4
        opt.zero_grad()
        loss.backward(inputs=opt.params)
        opt.step_dev()
        out['loss'] = loss.detach()

    for _ in range(3):
        opt.begin_step()
        step()
    torch.cuda.synchronize()
    mode = "one CUDA graph per step"
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            step()
        run = g.replay
    except Exception as e:                       # noqa: BLE001 — report, fall back to eager launches
        mode = "eager (capture failed: %s)" % str(e).splitlines()[0][:80]
        torch.cuda.synchronize()
        run = step
    for _ in range(3):
        opt.begin_step()
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        opt.begin_step()
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    return {"config": 3, "workload": "vgg16 Faster R-CNN train step (fwd + bwd + Adam), bs 1, 1x3x512x1024, 20 GT boxes, 1xB200",
            "ms_per_step": ms, "images_per_s": 1e3 / ms, "steps": steps, "execution": mode, "dtype": "bf16",
            "loss": float(out['loss'])}


def config5(dev):
    """NMS 1k-100k + IoU 30 720 x G"""
    from scda_b200 import _lib, synthetic
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    flush = _flush_buf(dev)
    ref = None
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libscda_ref.so")
    if os.path.exists(ref_path):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        try:
            import _reflib
            ref = _reflib.load()
        except Exception:                        # noqa: BLE001
            ref = None
    rows = []
    for n in (1000, 2000, 6000, 12000, 30720, 100000):
        boxes = torch.from_numpy(synthetic.nms_boxes(n, n)).to(dev)
        keep = torch.empty(n, dtype=torch.int64, device=dev)
        num = torch.zeros(1, dtype=torch.int64, device=dev)
        wsb = lib.scda_nms_workspace_bytes(n)
        ws = torch.empty(wsb // 8 + 1, dtype=torch.int64, device=dev)
        t = _time(lambda: lib.scda_nms(n, boxes.data_ptr(), 0.7, 0, keep.data_ptr(), num.data_ptr(), ws.data_ptr(),
                                       wsb, st), flush, iters=7 if n > 30000 else 10)
        row = {"op": "nms", "boxes": n, "us": round(t, 1), "kept": int(num.item()), "boxes_per_s": n / t * 1e6,
               "pair_tests_per_s": n * (n - 1) / 2 / t * 1e6}
        if ref is not None and n <= 30720:
            blocks = (n + 63) // 64
            mask = torch.empty(n * blocks, dtype=torch.int64, device=dev)
            pinned = torch.empty(n * blocks, dtype=torch.int64).pin_memory()

            def ref_pipeline():
                ref._nms(n, boxes.data_ptr(), mask.data_ptr(), 0.7)           # the reference's mask kernel …
                pinned.copy_(mask, non_blocking=True)                         # … + its bitmask D2H (host scan not timed)
            row["ref_kernel_plus_d2h_us"] = round(_time(ref_pipeline, flush, iters=5), 1)
        rows.append(row)
    for g in (8, 32, 128):
        b1 = torch.from_numpy(synthetic.nms_boxes(30720, 1)[:, :4].copy()).to(dev)
        b2 = torch.from_numpy(synthetic.gt_boxes(g, 2)[:, :4].copy()).to(dev)
        o = torch.empty(30720, g, device=dev)
        t = _time(lambda: lib.IOUOverlap(b1.data_ptr(), b2.data_ptr(), 4, 30720, g, o.data_ptr(), st), flush)
        row = {"op": "IOUOverlap", "shape": "30720x%d" % g, "us": round(t, 2), "pairs_per_s": 30720 * g / t * 1e6}
        if ref is not None:
            row["ref_us"] = round(_time(lambda: ref.IOUOverlap(b1.data_ptr(), b2.data_ptr(), 4, 30720, g,
                                                                o.data_ptr(), st), flush), 2)
        rows.append(row)
    return {"config": 5, "workload": "NMS(0.7) of 1k-100k score-sorted boxes (device scan included) + IoU 30 720 x G, 1xB200",
            "rows": rows}


def run(which, dev=None):
    dev = dev or torch.device("cuda", 0)
    from scda_b200 import _lib
    _lib.load()
    return {2: config2, 3: config3, 5: config5}[which](dev)


if __name__ == "__main__":
    torch.cuda.set_device(0)
    for c in ([int(a) for a in sys.argv[1:]] or [2, 3, 5]):
        print(json.dumps(run(c)), flush=True)
