"""Golden vectors for the HOST-side plumbing, produced by importing the reference's own
Python (functions/*.py, utils/*.py) from /root/reference in this container.

The reference modules do not import as shipped on torch 2.x / numpy 2.x
(SURVEY.md §8c), so this script supplies the minimum shims and nothing else:
  - `np.float = float`                         (utils/anchor_helper.py:46,47,53)
  - a stub `extensions` package whose `_cython_bbox.cython_bbox` is the reference's own
    compiled .pyx (oracle/_ref) and whose `nms` is oracle.nms (the restated gpu_nms,
    itself pinned bit-exact to the reference kernel by tests/golden/nms_*.npz)
  - `Tensor.cuda()` -> identity                (functions/*: hard-coded .cuda())
The numpy global RNG is seeded before each call; the oracle restatement, driven with
the same seed, must consume it identically.

Run here (needs /root/reference):  python tests/golden/make_golden_host.py
"""
import glob
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SCDA_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs  # noqa: E402
import oracle  # noqa: E402


def install_shims():
    np.float = float
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "cython_bbox*.so"))[0]
    spec = importlib.util.spec_from_file_location("cython_bbox", so)
    cy = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cy)
    ext = types.ModuleType("extensions")
    ext.__path__ = []
    sub = types.ModuleType("extensions._cython_bbox")
    sub.cython_bbox = cy
    ext._cython_bbox = sub
    ext.nms = lambda dets, thresh: torch.from_numpy(oracle.nms(dets.numpy(), thresh))
    ext.RoIPool = object
    sys.modules["extensions"] = ext
    sys.modules["extensions._cython_bbox"] = sub
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)


def cfg():
    c = json.load(open(os.path.join(REF, "examples/faster-rcnn/cityscapes/vgg/config_512.json")))
    for k, v in c.items():
        if k != "shared":
            v.update(c["shared"])
    return c


def main():
    install_shims()
    from utils import anchor_helper
    from functions.rpn_proposal import compute_rpn_proposals
    from functions.anchor_target import compute_anchor_targets
    from functions.proposal_target import compute_proposal_targets
    from functions.predict_bbox import compute_predicted_bboxes
    c = cfg()
    sh = c["shared"]
    out = {}
    out["anchors_32x64"] = anchor_helper.get_anchors_over_plane(32, 64, sh["anchor_ratios"], sh["anchor_scales"], sh["anchor_stride"])
    out["anchors_grid"] = anchor_helper.get_anchors_over_grid(sh["anchor_ratios"], sh["anchor_scales"], sh["anchor_stride"])

    info = np.array([[512, 1024, 0.5]], dtype=np.float32)
    cls, loc = _inputs.synth_rpn_outputs(0)
    out["rpn_cls"], out["rpn_loc"], out["image_info"] = cls, loc, info
    for tag in ("train", "test"):
        pc = c[tag + "_rpn_proposal_cfg"]
        out["proposals_" + tag] = compute_rpn_proposals(torch.from_numpy(cls), torch.from_numpy(loc), pc, torch.from_numpy(info)).numpy()

    gts = _inputs.gt_boxes(20, 0)[None]
    out["gts"] = gts
    np.random.seed(123)
    ct, lt, lm, norm = compute_anchor_targets((1, 60, 32, 64), c["train_anchor_target_cfg"], torch.from_numpy(gts), torch.from_numpy(info))
    out["anchor_cls_targets"], out["anchor_loc_targets"], out["anchor_loc_masks"] = ct.numpy(), lt.numpy(), lm.numpy()
    out["anchor_normalizer"] = np.array(norm)

    # proposals for the target stage: jitter the GTs so that there are positives to sample
    r = np.random.RandomState(5)
    jit = np.repeat(gts[0, :, :4], 30, axis=0) + r.normal(0, 6, (600, 4)).astype(np.float32)
    props = np.vstack([out["proposals_train"][:1400, 1:5], jit]).astype(np.float32)
    props = np.hstack([np.zeros((len(props), 1), np.float32), props, np.zeros((len(props), 1), np.float32)])
    out["pt_proposals"] = props
    np.random.seed(321)
    rois, lab, t, w = compute_proposal_targets(torch.from_numpy(props.copy()), c["train_proposal_target_cfg"], torch.from_numpy(gts), torch.from_numpy(info))
    out["pt_rois"], out["pt_labels"], out["pt_loc_targets"], out["pt_loc_weights"] = rois.numpy(), lab.numpy(), t.numpy(), w.numpy()

    # eval path
    rois_t = out["proposals_test"][:, :5].copy()
    r = np.random.RandomState(6)
    pc = torch.softmax(torch.from_numpy(r.standard_normal((len(rois_t), 9)).astype(np.float32) * 2), 1).numpy()
    pl = (r.standard_normal((len(rois_t), 36)) * 0.5).astype(np.float32)
    out["pb_cls"], out["pb_loc"] = pc, pl
    out["pb_out"] = compute_predicted_bboxes(torch.from_numpy(rois_t), torch.from_numpy(pc), torch.from_numpy(pl), info, c["test_predict_bbox_cfg"]).numpy()
    # fixtures stay small: the RPN maps are regenerated from the seed by the tests
    del out["rpn_cls"], out["rpn_loc"]
    np.savez_compressed(os.path.join(HERE, "host_plumbing.npz"), **out)
    json.dump(c, open(os.path.join(HERE, "..", "..", "scda_b200", "configs", "config_512_merged.json"), "w"), indent=1)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
