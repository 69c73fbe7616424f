"""oracle/host.py (numpy restatement of the reference's host plumbing) against golden
vectors produced by the reference's OWN Python modules (tests/golden/make_golden_host.py)."""
import os

import numpy as np
import pytest

import _inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_plumbing.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def cfg():
    return _inputs.load_cfg()


@pytest.fixture(scope="module")
def host():
    from oracle import host
    return host


def test_anchors(g, cfg, host):
    sh = cfg["shared"]
    a = host.anchors_over_plane(32, 64, sh["anchor_ratios"], sh["anchor_scales"], sh["anchor_stride"])
    assert a.dtype == np.float64 and np.array_equal(a, g["anchors_32x64"])
    assert np.array_equal(host.anchors_over_grid(sh["anchor_ratios"], sh["anchor_scales"], 16), g["anchors_grid"])
    # quirk: the ratios argument is ignored by the reference
    assert np.array_equal(host.anchors_over_grid([1], sh["anchor_scales"], 16), g["anchors_grid"])


@pytest.mark.parametrize("tag", ["train", "test"])
def test_rpn_proposals(g, cfg, host, tag):
    cls, loc = _inputs.synth_rpn_outputs(0)
    out = host.compute_rpn_proposals(cls, loc, cfg[tag + "_rpn_proposal_cfg"], g["image_info"])
    assert out.dtype == np.float32 and np.array_equal(out, g["proposals_" + tag])


def test_anchor_targets(g, cfg, host):
    np.random.seed(123)
    ct, lt, lm, norm = host.compute_anchor_targets((1, 60, 32, 64), cfg["train_anchor_target_cfg"],
                                                   g["gts"], g["image_info"])
    assert np.array_equal(ct, g["anchor_cls_targets"])
    assert np.array_equal(lt, g["anchor_loc_targets"])
    assert np.array_equal(lm, g["anchor_loc_masks"])
    assert norm == int(g["anchor_normalizer"])
    assert (ct == 1).sum() <= 128 and (ct >= 0).sum() == 256


def test_proposal_targets(g, cfg, host):
    np.random.seed(321)
    rois, lab, t, w = host.compute_proposal_targets(g["pt_proposals"].copy(), cfg["train_proposal_target_cfg"],
                                                    g["gts"], g["image_info"])
    assert np.array_equal(rois, g["pt_rois"])
    assert np.array_equal(lab, g["pt_labels"])
    assert np.array_equal(t, g["pt_loc_targets"])
    assert np.array_equal(w, g["pt_loc_weights"])
    assert rois.shape == (512, 5) and (lab > 0).sum() == 128


def test_predicted_bboxes(g, cfg, host):
    rois = g["proposals_test"][:, :5].copy()
    out = host.compute_predicted_bboxes(rois, g["pb_cls"], g["pb_loc"], g["image_info"], cfg["test_predict_bbox_cfg"])
    assert np.array_equal(out, g["pb_out"])


def test_crop_corners(host):
    # tools/faster_rcnn_train_val.py:411-438 on a 1024 x 512 image, 256 crops
    c = host.get_corner_from_center([(10, 10), (1020, 500), (512.7, 256.2), (100, 400)], 256, 1024, 512)
    assert c == [[0, 0, 256, 256], [768, 256, 1024, 512], [384, 128, 640, 384], [0, 256, 256, 512]]
    for x1, y1, x2, y2 in c:
        assert x2 - x1 == 256 and y2 - y1 == 256


def test_losses_against_torch(host):
    import torch
    import torch.nn.functional as F
    r = np.random.RandomState(0)
    x = r.standard_normal((300, 9)).astype(np.float32)
    t = r.randint(-1, 9, 300)
    ref = F.cross_entropy(torch.from_numpy(x), torch.from_numpy(t), ignore_index=-1).item()
    assert abs(host.cross_entropy(x, t) - ref) < 1e-5
    p = r.standard_normal((64, 36)).astype(np.float32) * 0.3
    q = r.standard_normal((64, 36)).astype(np.float32) * 0.3
    d = torch.from_numpy(p) - torch.from_numpy(q)
    a = d.abs()
    s = (a < 1 / 9.).float()
    ref = (d.pow(2) * 9 / 2. * s + (a - 0.5 / 9.) * (1 - s)).sum().item()
    assert abs(host.smooth_l1_loss_with_sigma(p, q) - ref) < 1e-3
