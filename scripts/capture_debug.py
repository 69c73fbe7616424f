#!/usr/bin/env python
"""Bisect which stage of the iteration breaks a CUDA-graph capture (debug aid)."""
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def main():
    stage = sys.argv[1]
    from scda_b200 import engine
    torch.cuda.set_device(0)
    cfg = bench.load_cfg()
    tr = engine.build_trainer(cfg, world_size=1, seed=0, use_graphs=False)
    image, target, gts, info = bench.synth_batch(0, pinned=False)
    image, target, gts = image.cuda(), target.cuda(), gts.cuda()
    t0 = time.perf_counter()
    tr.iteration(cfg, image, info, gts, target)
    torch.cuda.synchronize()
    print("eager iteration ok %.2fs" % (time.perf_counter() - t0), flush=True)
    m = tr.model
    x = {'cfg': cfg, 'image': image, 'image_info': info, 'ground_truth_bboxes': gts,
         'ignore_regions': None, 'cluster_num': 4, 'threshold': 128, 'device_clusters': True}

    def backbone():
        return m.feature_extractor(image)

    def rpn():
        return m.rpn(m.feature_extractor(image))

    def anchor():
        f = m.feature_extractor(image)
        c, l = m.rpn(f)
        fn = m._pin_args_to_fn(cfg, gts, info, None)
        return m._add_rpn_loss(fn['anchor_target_fn'], c, l)

    def props():
        from scda_b200.functions.rpn_proposal import rpn_proposals_device
        f = m.feature_extractor(image)
        c, l = m.rpn(f)
        return rpn_proposals_device(m._rpn_scores(c).data, l.data, cfg['train_rpn_proposal_cfg'], info)

    def forward():
        return m(x, target)

    def fwd_bwd():
        out = m(x, target)
        tr.opt.zero_grad()
        sum(out['losses']).backward(inputs=tr.opt.params)

    def whole():
        tr._static, tr._st = tr._by_shape[list(tr._by_shape)[0]]['static'], {}
        tr._body(tr._reduce_fn())

    fn = {'backbone': backbone, 'rpn': rpn, 'anchor': anchor, 'props': props, 'forward': forward,
          'fwd_bwd': fwd_bwd, 'whole': whole}[stage]
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    t0 = time.perf_counter()
    try:
        with torch.cuda.graph(g):
            fn()
        print("capture %s ok %.2fs" % (stage, time.perf_counter() - t0), flush=True)
        t0 = time.perf_counter()
        g.replay()
        torch.cuda.synchronize()
        print("replay %s ok %.4fs" % (stage, time.perf_counter() - t0), flush=True)
    except Exception:
        print("capture %s FAILED after %.2fs" % (stage, time.perf_counter() - t0), flush=True)
        traceback.print_exc(limit=6)


if __name__ == "__main__":
    main()
