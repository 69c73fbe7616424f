#!/usr/bin/env python
"""Wall time (with a device sync at each boundary) of the stages of one training iteration:
where the step is host-bound and where it is GPU-bound."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


class Tick(object):
    def __init__(self):
        self.t = {}
        self.last = None

    def start(self):
        torch.cuda.synchronize()
        self.last = time.perf_counter()

    def lap(self, name):
        cpu = time.perf_counter() - self.last
        torch.cuda.synchronize()
        tot = time.perf_counter() - self.last
        a = self.t.setdefault(name, [0.0, 0.0, 0])
        a[0] += cpu
        a[1] += tot
        a[2] += 1
        self.last = time.perf_counter()


def main():
    from scda_b200 import engine
    from scda_b200.models.faster_rcnn import faster_rcnn_adver_expansion_reweight_cluster as M
    torch.cuda.set_device(0)
    cfg = bench.load_cfg()
    tr = engine.build_trainer(cfg, world_size=1, seed=0, use_graphs=False, overlap=False)
    image, target, gts, info = bench.synth_batch(0, pinned=False)
    image, target, gts = image.cuda(), target.cuda(), gts.cuda()
    for _ in range(4):
        tr.iteration(cfg, image, info, gts, target)
    tk = Tick()

    # wrap the stages
    model = tr.model
    def wrap(obj, name, label):
        fn = getattr(obj, name)
        def w(*a, **k):
            tk.lap("(gap before %s)" % label)
            r = fn(*a, **k)
            tk.lap(label)
            return r
        setattr(obj, name, w)
    wrap(model, "feature_extractor", "backbone fwd")
    wrap(model, "rpn", "rpn head fwd")
    wrap(model, "rcnn", "rcnn head fwd")
    wrap(model, "_add_rpn_loss", "anchor targets + rpn loss")
    wrap(model, "_train_rois", "proposal targets")
    wrap(M, "rpn_proposals_device", "rpn proposals (+nms)")
    wrap(M, "cluster_targets_device", "cluster targets (k-means)")
    wrap(tr, "_seg_dis", "gan (1) dis")
    wrap(tr, "_seg_dis_patch", "gan (2) dis_patch")
    wrap(tr, "_seg_dec", "gan (3) dec")
    wrap(tr, "_seg_fake", "gan (4) fake fwd")
    wrap(tr.opt, "zero_grad", "det zero_grad")
    wrap(tr.opt, "step_dev", "det adam (after backward)")
    n = 5
    t0 = time.perf_counter()
    for _ in range(n):
        tk.start()
        tr.iteration(cfg, image, info, gts, target)
        tk.lap("(tail)")
    print("instrumented iteration: %.2f ms" % ((time.perf_counter() - t0) / n * 1e3))
    print("%-36s %9s %9s %6s" % ("stage", "host ms", "total ms", "calls"))
    for k, (c, t, m) in tk.t.items():
        print("%-36s %9.3f %9.3f %6d" % (k, c / n * 1e3, t / n * 1e3, m // n))


if __name__ == "__main__":
    main()
