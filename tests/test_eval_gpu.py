"""Evaluation path and input pipeline on the GPU: the batched per-class NMS against the single-problem NMS
(itself bit-exact against the reference kernel, tests/test_ops_gpu.py), compute_predicted_bboxes against the
oracle's numpy restatement in both row orders, the image-preparation kernel against PIL-style nearest resize /
torch bilinear + ToTensor + Normalize."""
import numpy as np
import pytest

import _inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("groups,n_cap,seed", [(8, 300, 0), (3, 1024, 1), (16, 65, 2), (1, 1, 3)])
def test_nms_groups_equals_single_nms(cuda_lib, groups, n_cap, seed):
    import torch
    from scda_b200.extensions._nms.pth_nms import nms_device
    from scda_b200.functions.predict_bbox import nms_groups
    r = np.random.RandomState(seed)
    dets = np.zeros((groups, n_cap, 5), np.float32)
    n_live = r.randint(0, n_cap + 1, groups).astype(np.int32)
    n_live[0] = n_cap
    for g in range(groups):
        b = _inputs.clustered_boxes(n_cap, seed * 100 + g) if g % 2 else _inputs.nms_boxes(n_cap, seed * 100 + g)
        dets[g] = b
    d = torch.from_numpy(dets).cuda()
    keep, num = nms_groups(d, torch.from_numpy(n_live).cuda(), 0.5)
    for g in range(groups):
        n = int(n_live[g])
        if n == 0:
            assert int(num[g]) == 0
            continue
        k_ref, n_ref = nms_device(d[g, :n].contiguous(), 0.5)
        assert int(num[g]) == int(n_ref)
        assert torch.equal(keep[g, :int(n_ref)], k_ref[:int(n_ref)])


@pytest.mark.parametrize("top_n", [100, -1])
def test_predicted_bboxes_vs_oracle(cuda_lib, top_n):
    import torch
    from oracle import host
    from scda_b200.functions.predict_bbox import compute_predicted_bboxes
    cfg = dict(_inputs.load_cfg()["test_predict_bbox_cfg"], top_n=top_n)
    r = np.random.RandomState(5)
    B, n = 2, 150
    rois = np.concatenate([_inputs.rois_uniform(n, 10 + b, img_w=1024, img_h=512, wh=(16, 300)) for b in range(B)])
    rois[:n, 0], rois[n:, 0] = 0, 1
    cls = r.dirichlet(np.ones(9) * 0.3, B * n).astype(np.float32)
    loc = (r.standard_normal((B * n, 36)) * 0.5).astype(np.float32)
    info = np.array([[512, 1024, 1.0], [512, 1024, 1.0]], np.float32)
    ref = host.compute_predicted_bboxes(rois, cls, loc, info, cfg)
    out = compute_predicted_bboxes(torch.from_numpy(rois).cuda(), torch.from_numpy(cls).cuda(),
                                   torch.from_numpy(loc).cuda(), info, cfg).cpu().numpy()
    assert out.shape == ref.shape
    # same rows in the same order (scores are distinct): batch, score and class exact, corners through exp()
    assert np.array_equal(out[:, 0], ref[:, 0]) and np.array_equal(out[:, 6], ref[:, 6])
    np.testing.assert_allclose(out[:, 5], ref[:, 5], rtol=0, atol=0)
    np.testing.assert_allclose(out[:, 1:5], ref[:, 1:5], rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("n,thresh,seed", [(300, 0.0, 0), (300, 0.05, 1), (1024, 0.0, 2), (37, 0.3, 3), (1, 0.0, 4)])
def test_predicted_bboxes_single_image_kernels(cuda_lib, n, thresh, seed):
    """One image: the three-launch form (csrc/predict_ops.cu around scda_nms_groups) against the oracle's numpy
    restatement of functions/predict_bbox.py:13-66 and, row for row, against the tensor-op form."""
    import torch
    from oracle import host
    from scda_b200.functions import predict_bbox as pb
    cfg = dict(_inputs.load_cfg()["test_predict_bbox_cfg"], top_n=100, score_thresh=thresh)
    r = np.random.RandomState(seed)
    rois = _inputs.rois_uniform(n, 20 + seed, img_w=1024, img_h=512, wh=(16, 300))
    rois[:, 0] = 0
    cls = r.dirichlet(np.ones(9) * 0.3, n).astype(np.float32)
    loc = (r.standard_normal((n, 36)) * 0.5).astype(np.float32)
    info = np.array([[512, 1024, 1.0]], np.float32)
    args = (torch.from_numpy(rois).cuda(), torch.from_numpy(cls).cuda(), torch.from_numpy(loc).cuda(), info, cfg)
    out = pb.compute_predicted_bboxes(*args)
    try:
        pb.FORCE_TENSOR_PATH = True
        slow = pb.compute_predicted_bboxes(*args)
    finally:
        pb.FORCE_TENSOR_PATH = False
    assert out.shape == slow.shape and torch.equal(out, slow)
    ref = host.compute_predicted_bboxes(rois, cls, loc, info, cfg)
    out = out.cpu().numpy()
    assert out.shape == ref.shape
    assert np.array_equal(out[:, 0], ref[:, 0]) and np.array_equal(out[:, 6], ref[:, 6])
    np.testing.assert_allclose(out[:, 5], ref[:, 5], rtol=0, atol=0)
    np.testing.assert_allclose(out[:, 1:5], ref[:, 1:5], rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("mode,flip,shape,new", [("nearest", False, (1024, 2048), (512, 1024)),
                                                 ("nearest", True, (375, 500), (600, 800)),
                                                 ("bilinear", False, (1024, 2048), (512, 1024)),
                                                 ("bilinear", True, (300, 451), (512, 770))])
def test_image_prepare(cuda_lib, mode, flip, shape, new):
    import torch
    import torch.nn.functional as F
    from scda_b200.datasets.example_dataset import prepare_image
    r = np.random.RandomState(0)
    img = r.randint(0, 256, shape + (3,), dtype=np.uint8)
    out = prepare_image(img, new[0], new[1], flip=flip, mode=mode).cpu()
    assert out.shape == (1, 3) + new and out.dtype == torch.float32
    src = torch.from_numpy(img).permute(2, 0, 1).float()[None]
    if mode == "nearest":
        # Pillow < 7 Image.resize((w, h)) default: source index floor((dst + 0.5) * scale)
        ys = np.minimum(np.floor((np.arange(new[0]) + 0.5) * shape[0] / new[0]).astype(np.int64), shape[0] - 1)
        xs = np.minimum(np.floor((np.arange(new[1]) + 0.5) * shape[1] / new[1]).astype(np.int64), shape[1] - 1)
        ref = src[:, :, torch.from_numpy(ys)][:, :, :, torch.from_numpy(xs)]
    else:
        ref = F.interpolate(src, size=new, mode="bilinear", align_corners=False)
    if flip:
        ref = ref.flip(3)
    ref = (ref / 255.0 - 0.5) / 0.5
    tol = 1e-6 if mode == "nearest" else 2e-5
    assert float((out - ref).abs().max()) <= tol


def test_collate_pads_like_the_reference(cuda_lib):
    import torch
    from scda_b200.datasets.example_dataset import collate
    r = np.random.RandomState(1)
    a = r.randint(0, 256, (100, 200, 3), dtype=np.uint8)
    b = r.randint(0, 256, (80, 120, 3), dtype=np.uint8)
    boxes_a = np.array([[1, 2, 30, 40, 3], [5, 6, 70, 80, 1]], np.float32)
    boxes_b = np.array([[2, 2, 20, 20, 2]], np.float32)
    ig = np.zeros((1, 4), np.float32)
    batch = [(a, (50, 100, 0.5, False, boxes_a, ig, "a.png")), (b, (40, 60, 0.5, True, boxes_b, ig, "b.png"))]
    images, info, gts, igs, names = collate(batch)
    assert images.shape == (2, 3, 50, 100) and images.is_cuda
    assert torch.equal(info, torch.tensor([[50, 100, 0.5], [40, 60, 0.5]]))
    assert gts.shape == (2, 2, 5) and torch.equal(gts[1, 1], torch.zeros(5)) and names == ["a.png", "b.png"]
    assert float(images[1, :, 40:, :].abs().max()) == 0 and float(images[1, :, :, 60:].abs().max()) == 0
    assert float(images[1, :, :40, :60].abs().max()) > 0
