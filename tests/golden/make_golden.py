"""Generate the golden vectors under tests/golden/ from the REFERENCE itself.

Two sources, both the reference's own code (oracle/build.py compiles them in
place from /root/reference; nothing is copied):
  - oracle/_ref/cython_bbox*.so  : the reference's cython_bbox.pyx -> runs on any CPU
  - oracle/_ref/libscda_ref.so   : the reference's six .cu files, unmodified, sm_100a
                                   -> needs a GPU; run on the B200 box:
        gpurun -- 'python tests/golden/make_golden.py --gpu --out gpurun_out/golden'
    and copy gpurun_out/golden/*.npz here.

Inputs are the seeded generators of tests/_inputs.py at small sizes, so the
fixtures stay a few hundred KB.  The reference's gpu_nms host half
(nms_cuda.c:41-58) cannot be compiled (TH); the GPU fixture stores the
reference kernel's bitmask, and the kept indices derived from it by that
18-line scan restated in oracle.nms_scan.
"""
import argparse
import glob
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _inputs  # noqa: E402


def cpu_golden(out):
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "cython_bbox*.so"))
    if not so:
        print("cython_bbox reference not built; skipping")
        return
    spec = importlib.util.spec_from_file_location("cython_bbox", so[0])
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    boxes = _inputs.nms_boxes(700, 11)[:, :4].copy()
    query = _inputs.gt_boxes(20, 0)[:, :4].copy()
    # degenerate rows: zero-area, touching edges, identical boxes
    boxes[0] = query[0]
    boxes[1] = [10, 10, 10, 50]
    boxes[2] = [query[1][2], query[1][1], query[1][2] + 30, query[1][3]]
    np.savez_compressed(os.path.join(out, "cython_bbox_overlaps.npz"), boxes=boxes, query=query,
                        overlaps=m.bbox_overlaps(boxes, query))
    print("wrote cython_bbox_overlaps.npz")


def gpu_golden(out):
    import torch
    assert torch.cuda.is_available()
    import _gpu_ops as G
    import _reflib
    import oracle
    ref = _reflib.load()

    feat = _inputs.features((2, 16, 20, 24), 0)
    rois = _inputs.rois_uniform(24, 1, img_w=24 * 16, img_h=20 * 16, wh=(8, 300), batch=2)
    rois[0, 1:] = [-40, -40, 30, 30]          # partly outside
    rois[1, 1:] = [100, 100, 90, 90]          # malformed (end < start)
    rois[2, 1:] = [500, 500, 600, 600]        # wholly outside
    scale = 1.0 / 16
    po, pa = G.roi_pool_fwd(ref, feat, rois, 7, 7, scale)
    g = _inputs.features(po.shape, 2)
    pg = G.roi_pool_bwd(ref, g, rois, pa, feat.shape, scale)
    np.savez_compressed(os.path.join(out, "roi_pool.npz"), feat=feat, rois=rois, scale=scale,
                        out=po, argmax=pa, top_diff=g, bottom_diff=pg)

    ao = G.roi_align_fwd(ref, feat, rois, 8, 8, scale)
    ga = _inputs.features(ao.shape, 3)
    ag = G.roi_align_bwd(ref, ga, rois, feat.shape, scale)
    np.savez_compressed(os.path.join(out, "roi_align.npz"), feat=feat, rois=rois, scale=scale,
                        out=ao, top_diff=ga, bottom_diff=ag)

    for name, boxes, th in (("nms_uniform", _inputs.nms_boxes(1500, 5), 0.7),
                            ("nms_clustered", _inputs.clustered_boxes(1000, 6), 0.5)):
        mask = G.nms_mask(ref, boxes, th)
        keep = oracle.nms_scan(mask)
        np.savez_compressed(os.path.join(out, name + ".npz"), boxes=boxes, thresh=th,
                            keep=keep, mask=mask)

    b1 = _inputs.nms_boxes(300, 7)
    b2 = _inputs.gt_boxes(17, 8)
    np.savez_compressed(os.path.join(out, "iou_overlap.npz"), b1=b1, b2=b2,
                        out=G.iou_overlap(ref, b1, b2))

    x, t = _inputs.focal_inputs(500, 8, 9)
    l, dx = G.sigmoid_focal(ref, x, t, 37.0, 2.0, 0.25)
    np.savez_compressed(os.path.join(out, "sigmoid_focal.npz"), logits=x, targets=t,
                        weight_pos=37.0, gamma=2.0, alpha=0.25, losses=l, dx=dx)
    x, t = _inputs.focal_inputs(500, 9, 10, softmax=True)
    l, p, dx, buff = G.softmax_focal(ref, x, t, 37.0, 2.0, 0.25)
    np.savez_compressed(os.path.join(out, "softmax_focal.npz"), logits=x, targets=t,
                        weight_pos=37.0, gamma=2.0, alpha=0.25, losses=l, priors=p, dx=dx,
                        buff=buff)
    print("wrote GPU goldens to", out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--out", default=HERE)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    cpu_golden(a.out)
    if a.gpu:
        gpu_golden(a.out)
