"""The detector graph end to end on the GPU: one training forward/backward and one eval
forward through the reference-facing modules (models.faster_rcnn.*), plus the drop-in
check that code written against the reference's top-level package names imports ours."""
import numpy as np
import pytest

import _inputs

pytestmark = pytest.mark.gpu


def _batch(seed=0, h=256, w=512):
    import torch
    r = np.random.RandomState(seed)
    img = torch.from_numpy(r.standard_normal((1, 3, h, w)).astype(np.float32)).cuda()
    tgt = torch.from_numpy(r.standard_normal((1, 3, h, w)).astype(np.float32)).cuda()
    gts = torch.from_numpy(_inputs.gt_boxes(12, seed, img_w=w, img_h=h)[None])
    info = torch.tensor([[h, w, 0.5]])
    return img, tgt, gts, info


def test_reference_import_names_resolve(cuda_lib):
    from scda_b200 import compat
    compat.install()
    from extensions import nms, RoIPool                       # functions/rpn_proposal.py:4
    from extensions._cython_bbox import cython_bbox           # utils/bbox_helper.py:5
    from models.faster_rcnn.vgg_adver_expansion_cluster import vgg16 as VGG16   # tools/...:31
    from models.faster_rcnn.faster_rcnn_adver_expansion_reweight_cluster import (
        GAN_dis_AE_patch, GAN_dis_AE, GAN_decoder_AE)
    from functions.rpn_proposal import compute_rpn_proposals
    from functions.proposal_target import compute_proposal_targets
    from utils.distributed_utils import dist_init, average_gradients, broadcast_params
    assert callable(nms) and callable(VGG16) and callable(compute_rpn_proposals)


def test_detector_train_forward_backward(cuda_lib):
    import torch
    from scda_b200.models.faster_rcnn.vgg_adver_expansion_cluster import vgg16
    cfg = _inputs.load_cfg()
    torch.manual_seed(0)
    np.random.seed(0)
    model = vgg16(cfg=cfg["shared"]).cuda().train()
    img, tgt, gts, info = _batch()
    x = {"cfg": cfg, "image": img, "image_info": info, "ground_truth_bboxes": gts,
         "ignore_regions": None, "cluster_num": 4, "threshold": 128}
    out = model(x, tgt)
    assert len(out["losses"]) == 4 and all(torch.isfinite(l) for l in out["losses"])
    src, dst = out["cluster_features"]
    assert src.shape == (4, 128, 4096) and dst.shape == (4, 128, 4096) and not src.requires_grad
    assert np.asarray(out["cluster_centers"][0]).shape == (4, 2)
    loss = sum(out["losses"])
    loss.backward()
    g = model.features[0].weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0
    assert model.fc_rcnn_loc.weight.grad is not None
    assert 0.0 <= float(out["accuracy"][0]) <= 100.0


def test_detector_eval_forward(cuda_lib):
    import torch
    from scda_b200.models.faster_rcnn.vgg_adver_expansion_cluster import vgg16
    cfg = _inputs.load_cfg()
    torch.manual_seed(0)
    model = vgg16(cfg=cfg["shared"]).cuda().eval()
    img, _, _, info = _batch()
    x = {"cfg": cfg, "image": img, "image_info": info, "ground_truth_bboxes": None,
         "ignore_regions": None}
    with torch.no_grad():
        out = model(x)
    proposals, bboxes = out["predict"]
    assert proposals.shape[1] == 5 and proposals.shape[0] <= 300
    assert bboxes.shape[1] == 7 and bboxes.shape[0] <= 100
    assert bboxes.is_cuda
