#!/usr/bin/env python
"""Which stage of the training iteration owns which GPU kernels: one eager iteration under
torch.profiler with a record_function label per stage; every kernel is charged to the innermost
stage label above the op that launched it (backward ops are charged to "backward:<phase>").
Prints per stage: kernel count, GPU time, and the top kernels."""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def main():
    from torch.profiler import ProfilerActivity, profile, record_function
    from scda_b200 import engine
    from scda_b200.models.faster_rcnn import faster_rcnn_adver_expansion_reweight_cluster as M
    torch.cuda.set_device(0)
    cfg = bench.load_cfg()
    tr = engine.build_trainer(cfg, world_size=1, seed=0, use_graphs=False, overlap=False)
    image, target, gts, info = bench.synth_batch(0, pinned=False)
    image, target, gts = image.cuda(), target.cuda(), gts.cuda()
    for _ in range(3):
        tr.iteration(cfg, image, info, gts, target)

    def wrap(obj, name, label):
        fn = getattr(obj, name)

        def w(*a, **k):
            with record_function("stage:" + label):
                return fn(*a, **k)
        setattr(obj, name, w)
    model = tr.model
    wrap(model, "feature_extractor", "backbone fwd")
    wrap(model, "rpn", "rpn head fwd")
    wrap(model, "rcnn", "rcnn head fwd")
    wrap(model, "_add_rpn_loss", "anchor targets + rpn loss")
    wrap(model, "_add_rcnn_loss", "rcnn loss")
    wrap(model, "_train_rois", "proposal targets")
    wrap(M, "rpn_proposals_device", "rpn proposals (+nms)")
    wrap(M, "cluster_targets_device", "cluster targets (k-means)")
    wrap(engine, "crops_device", "crops")
    wrap(tr, "_seg_dis", "gan (1) dis")
    wrap(tr, "_seg_dis_patch", "gan (2) dis_patch")
    wrap(tr, "_seg_dec", "gan (3) dec")
    wrap(tr, "_seg_fake", "gan (4) fake fwd")
    wrap(tr, "_seg_det_backward", "detector loss+backward")
    for o, n in ((tr.opt, "det"), (tr.opt_dec, "dec"), (tr.opt_dis, "dis"), (tr.opt_dis_patch, "dis_patch")):
        wrap(o, "step_dev", "adam " + n)
        wrap(o, "zero_grad", "zero_grad " + n)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        tr.iteration(cfg, image, info, gts, target)
        torch.cuda.synchronize()

    evs = prof.events()
    stage_t = collections.Counter()
    stage_n = collections.Counter()
    per = collections.defaultdict(collections.Counter)
    for e in evs:
        if not e.kernels:
            continue
        # innermost op only (children also list the kernels): skip events whose child owns them
        if any(c.kernels for c in e.cpu_children):
            continue
        label = None
        p = e
        in_bwd = False
        while p is not None:
            if p.name.startswith("stage:"):
                label = p.name[6:]
                break
            if "Backward" in p.name or p.name.startswith("autograd::engine"):
                in_bwd = True
            p = p.cpu_parent
        if label is None:
            label = "(backward thread)" if in_bwd or e.thread != evs[0].thread else "(other)"
        for k in e.kernels:
            stage_t[label] += k.duration
            stage_n[label] += 1
            per[label][k.name[:70]] += k.duration
    tot = sum(stage_t.values())
    print("total kernel time %.1f us, %d kernels" % (tot, sum(stage_n.values())))
    for label, t in stage_t.most_common():
        print("%-34s %8.1f us %5d kernels" % (label, t, stage_n[label]))
        for name, d in per[label].most_common(6):
            print("      %8.1f  %s" % (d, name))
    # backward-thread kernels by autograd node
    node_t = collections.Counter()
    node_n = collections.Counter()
    for e in evs:
        if not e.kernels or any(c.kernels for c in e.cpu_children):
            continue
        p, top = e, None
        while p is not None:
            if p.name.startswith("stage:"):
                top = None
                break
            if "Backward" in p.name:
                top = p.name
            p = p.cpu_parent
        if top:
            for k in e.kernels:
                node_t[top[:60]] += k.duration
                node_n[top[:60]] += 1
    print("\nbackward-thread kernels by autograd node")
    for name, t in node_t.most_common(40):
        print("%-62s %8.1f us %5d" % (name, t, node_n[name]))


if __name__ == "__main__":
    main()
