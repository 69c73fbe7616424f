"""Parity tests proper (GPU): libscda_b200 through its C ABI against
  (1) the CPU oracle on the same seeded inputs,
  (2) the reference's own kernels compiled unmodified (oracle/_ref/libscda_ref.so)
      driven through the SAME binding (tests/_gpu_ops.py),
  (3) size-independent properties at BASELINE.json's full sizes.
Bars: bit-exact for RoIPool output/argmax, NMS bitmask and kept indices and
both IoU matrices; rel 1e-4 (+ tiny abs) for RoIAlign and the focal losses."""
import numpy as np
import pytest

import _gpu_ops as G
import _inputs
import _reflib

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-6   # north_star: 1e-4 rel for fp32 RoIAlign / losses


@pytest.fixture(scope="module")
def ref_lib():
    if not _reflib.available():
        pytest.skip("oracle/_ref/libscda_ref.so not built")
    return _reflib.load()


def upper_words(mask):
    n, cb = mask.shape
    rb = (np.arange(n) // 64)[:, None]
    return np.where(np.arange(cb)[None, :] >= rb, mask, 0)


POOL_CASES = [
    # (feat shape, n rois, pooled, img_w, img_h, wh range)
    ((1, 8, 32, 64), 64, (7, 7), 1024, 512, (16, 512)),
    ((2, 12, 20, 24), 33, (7, 7), 384, 320, (4, 300)),
    ((1, 6, 9, 13), 17, (3, 5), 208, 144, (2, 150)),       # C % 4 != 0 -> scalar store path
    ((1, 256, 64, 64), 128, (7, 7), 1024, 1024, (16, 512)),  # config 1 shape
    ((1, 8, 21, 70), 40, (7, 7), 1120, 336, (8, 1100)),    # W not a multiple of 32, wide RoIs
    ((1, 4, 40, 300), 24, (7, 9), 4800, 640, (8, 4800)),   # very wide map
    ((2, 8, 16, 16), 30, (12, 3), 256, 256, (4, 250)),     # pooled_height > 8
    ((1, 512, 32, 64), 512, (7, 7), 1024, 512, (16, 512)),   # the model's operating point (plane-resident form)
    ((1, 2, 200, 300), 12, (7, 7), 4800, 3200, (8, 4000)),   # plane too large for shared memory: warp form
    ((1, 2, 200, 300), 12, (9, 7), 4800, 3200, (8, 4000)),   # ... and pooled_height > 8: per-output form
    ((3, 5, 12, 20), 50, (7, 7), 320, 192, (4, 300)),        # three images interleaved in the RoI list
]


@pytest.mark.parametrize("shape,R,pool,iw,ih,wh", POOL_CASES)
def test_roi_pool_bit_exact(cuda_lib, ref_lib, oracle_mod, shape, R, pool, iw, ih, wh):
    feat = _inputs.features(shape, 0)
    rois = _inputs.rois_uniform(R, 1, img_w=iw, img_h=ih, wh=wh, batch=shape[0])
    rois[0, 1:] = [-40, -40, 30, 30]
    rois[1, 1:] = [100, 100, 90, 90]
    rois[2, 1:] = [iw + 200, ih + 200, iw + 300, ih + 300]
    scale = 1 / 16.
    out, arg = G.roi_pool_fwd(cuda_lib, feat, rois, pool[0], pool[1], scale)
    o_or, a_or = oracle_mod.roi_pool_forward(feat, rois, pool[0], pool[1], scale)
    o_rf, a_rf = G.roi_pool_fwd(ref_lib, feat, rois, pool[0], pool[1], scale)
    assert np.array_equal(out, o_or) and np.array_equal(arg, a_or)
    assert np.array_equal(out, o_rf) and np.array_equal(arg, a_rf)

    g = _inputs.features(out.shape, 2)
    gi = G.roi_pool_bwd(cuda_lib, g, rois, arg, feat.shape, scale)
    gi_rf = G.roi_pool_bwd(ref_lib, g, rois, arg, feat.shape, scale)
    gi_or = oracle_mod.roi_pool_backward(g, rois, arg, feat.shape, scale)
    assert np.array_equal(gi_rf, gi_or)              # oracle reproduces the reference's order
    # scatter order differs: sums of up to hundreds of terms, rel 1e-4 of the plane's scale
    assert np.abs(gi - gi_rf).max() <= RTOL * max(1.0, float(np.abs(gi_rf).max()))


def test_roi_pool_without_argmax_and_empty(cuda_lib):
    import torch
    feat = torch.randn(1, 8, 16, 16, device="cuda")
    rois = torch.tensor([[0, 0, 0, 100, 100]], dtype=torch.float32, device="cuda")
    out = torch.zeros(1, 8, 7, 7, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    assert cuda_lib.ROIPoolForwardLaucher(feat.data_ptr(), 1 / 16., 1, 16, 16, 8, 7, 7,
                                          rois.data_ptr(), out.data_ptr(), None, s) == 1
    assert cuda_lib.ROIPoolForwardLaucher(feat.data_ptr(), 1 / 16., 0, 16, 16, 8, 7, 7,
                                          rois.data_ptr(), out.data_ptr(), None, s) == 1
    # rejected arguments return 0 instead of exit(-1)
    assert cuda_lib.ROIPoolForwardLaucher(feat.data_ptr(), 1 / 16., 1, 16, 16, 8, 0, 7,
                                          rois.data_ptr(), out.data_ptr(), None, s) == 0
    torch.cuda.synchronize()


@pytest.mark.parametrize("shape,R,al,iw,ih,wh", [
    ((1, 256, 64, 64), 128, (7, 7), 1024, 1024, (16, 512)),   # BASELINE.json configs[0]
    ((1, 256, 64, 64), 128, (8, 8), 1024, 1024, (16, 512)),   # the RoIAlignAvg/Max intermediate
    ((2, 10, 20, 24), 33, (5, 3), 384, 320, (4, 300)),
    ((3, 6, 13, 21), 40, (7, 7), 336, 208, (4, 300)),         # W % 4 != 0: scalar reductions in the backward
    ((1, 2, 200, 300), 12, (7, 7), 4800, 3200, (8, 4000)),    # plane too large for shared memory: per-RoI form
])
def test_roi_align_parity(cuda_lib, ref_lib, oracle_mod, shape, R, al, iw, ih, wh):
    feat = _inputs.features(shape, 0)
    rois = _inputs.rois_uniform(R, 1, img_w=iw, img_h=ih, wh=wh, batch=shape[0])
    rois[0, 1:] = [-40, -40, 30, 30]
    rois[1, 1:] = [100, 100, 90, 90]
    scale = 1 / 16.
    out = G.roi_align_fwd(cuda_lib, feat, rois, al[0], al[1], scale)
    o_rf = G.roi_align_fwd(ref_lib, feat, rois, al[0], al[1], scale)
    o_or = oracle_mod.roi_align_forward(feat, rois, al[0], al[1], scale)
    np.testing.assert_allclose(out, o_rf, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out, o_or, rtol=RTOL, atol=ATOL)
    assert np.mean(out == o_rf) > 0.999          # in fact the same bits almost everywhere

    g = _inputs.features(out.shape, 1)
    gi = G.roi_align_bwd(cuda_lib, g, rois, feat.shape, scale)
    gi_rf = G.roi_align_bwd(ref_lib, g, rois, feat.shape, scale)
    gi_or = oracle_mod.roi_align_backward(g, rois, feat.shape, scale)
    # sums of up to hundreds of terms in different orders: rel 1e-4 of the plane's scale
    tol = RTOL * max(1.0, float(np.abs(gi_or).max()))
    assert np.abs(gi - gi_rf).max() <= tol
    assert np.abs(gi - gi_or).max() <= tol


NMS_CASES = [(1, 0, 0.7, "u"), (64, 1, 0.7, "u"), (65, 2, 0.5, "c"), (1000, 3, 0.7, "u"),
             (2000, 4, 0.5, "c"), (6000, 6000, 0.7, "u"), (12000, 12000, 0.7, "u"),
             (12000, 7, 0.5, "c")]


@pytest.mark.parametrize("n,seed,th,kind", NMS_CASES)
def test_nms_bit_exact(cuda_lib, ref_lib, oracle_mod, n, seed, th, kind):
    import torch
    boxes = _inputs.nms_boxes(n, seed) if kind == "u" else _inputs.clustered_boxes(n, seed)
    m_rf = G.nms_mask(ref_lib, boxes, th)
    m_us = G.nms_mask(cuda_lib, boxes, th)
    assert np.array_equal(upper_words(m_us), upper_words(m_rf))
    keep_ref = oracle_mod.nms_scan(m_rf)             # reference kernel + restated host scan
    keep_or = oracle_mod.nms(boxes, th)              # pure CPU restatement
    assert np.array_equal(keep_ref, keep_or)

    b = torch.from_numpy(boxes).cuda()
    keep = torch.full((n,), -1, dtype=torch.int64, device="cuda")
    num = torch.zeros(1, dtype=torch.int64, device="cuda")
    ws_bytes = cuda_lib.scda_nms_workspace_bytes(n)
    ws = torch.empty(ws_bytes // 8 + 1, dtype=torch.int64, device="cuda")
    st = cuda_lib.scda_nms(n, b.data_ptr(), th, 0, keep.data_ptr(), num.data_ptr(),
                           ws.data_ptr(), ws_bytes, torch.cuda.current_stream().cuda_stream)
    assert st == 1
    k = int(num.item())
    assert np.array_equal(keep[:k].cpu().numpy(), keep_ref)

    # early stop: first max_keep survivors, identical prefix
    mk = max(1, len(keep_ref) // 3)
    st = cuda_lib.scda_nms(n, b.data_ptr(), th, mk, keep.data_ptr(), num.data_ptr(),
                           ws.data_ptr(), ws_bytes, torch.cuda.current_stream().cuda_stream)
    assert st == 1 and int(num.item()) == mk
    assert np.array_equal(keep[:mk].cpu().numpy(), keep_ref[:mk])


def test_nms_full_size_properties(cuda_lib, oracle_mod):
    """Config 5 upper sizes: idempotence + subset/order, and equality with the CPU
    restatement where it finishes in seconds."""
    import torch
    from scda_b200.extensions._nms.pth_nms import nms_device
    for n in (30720, 100000):
        boxes = _inputs.nms_boxes(n, n)
        d = torch.from_numpy(boxes).cuda()
        keep, num = nms_device(d, 0.7)
        k = keep[: int(num.item())]
        kn = k.cpu().numpy()
        assert np.all(np.diff(kn) > 0) and kn[0] == 0
        keep2, num2 = nms_device(d[k].contiguous(), 0.7)
        assert int(num2.item()) == len(kn)                       # idempotent
        if n <= 30720:
            assert np.array_equal(kn, oracle_mod.nms(boxes, 0.7))


def test_pth_nms_api(cuda_lib, oracle_mod):
    import torch
    from scda_b200.extensions import nms
    boxes = _inputs.nms_boxes(3000, 9)
    out = nms(torch.from_numpy(boxes), 0.7)                      # CPU tensor in, as rpn_proposal.py:64
    assert out.device.type == "cpu" and out.dtype == torch.int64
    assert np.array_equal(out.numpy(), oracle_mod.nms(boxes, 0.7))
    assert nms(torch.zeros(0, 5), 0.7).numel() == 0


@pytest.mark.parametrize("n1,n2", [(30720, 8), (30720, 32), (30720, 128), (2020, 20), (1, 1), (300, 700)])
def test_iou_matrices_bit_exact(cuda_lib, ref_lib, oracle_mod, n1, n2):
    import torch
    b1 = _inputs.nms_boxes(n1, 1)[:, :4].copy()
    b2 = _inputs.gt_boxes(n2, 2)[:, :4].copy()
    b1[0] = b2[0]
    assert np.array_equal(G.iou_overlap(cuda_lib, b1, b2), G.iou_overlap(ref_lib, b1, b2))
    assert np.array_equal(G.iou_overlap(cuda_lib, b1, b2), oracle_mod.iou_overlap(b1, b2))
    a, q = torch.from_numpy(b1).cuda(), torch.from_numpy(b2).cuda()
    out = torch.zeros(n1, n2, device="cuda")
    assert cuda_lib.scda_bbox_overlaps(n1, a.data_ptr(), n2, q.data_ptr(), out.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream) == 1
    assert np.array_equal(out.cpu().numpy(), oracle_mod.bbox_overlaps(b1, b2))


def test_cython_bbox_api_matches_golden(cuda_lib):
    import os
    from scda_b200.extensions._cython_bbox import cython_bbox
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cython_bbox_overlaps.npz"))
    assert np.array_equal(cython_bbox.bbox_overlaps(g["boxes"], g["query"]), g["overlaps"])


@pytest.mark.parametrize("m,k", [(500, 8), (30720, 8), (7, 1)])
def test_sigmoid_focal_parity(cuda_lib, ref_lib, oracle_mod, m, k):
    import torch
    x, t = _inputs.focal_inputs(m, k, 3)
    args = (23.0, 2.0, 0.25)
    l, dx = G.sigmoid_focal(cuda_lib, x, t, *args)
    l_rf, dx_rf = G.sigmoid_focal(ref_lib, x, t, *args)
    np.testing.assert_allclose(l, l_rf, rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(dx, dx_rf, rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(l, oracle_mod.sigmoid_focal_forward(x, t, *args), rtol=RTOL, atol=1e-8)
    np.testing.assert_allclose(dx, oracle_mod.sigmoid_focal_backward(x, t, *args), rtol=RTOL, atol=1e-8)
    # fused sum == sum of the reference's per-element losses
    xs, ts = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
    tot = torch.zeros(1, device="cuda")
    assert cuda_lib.scda_sigmoid_focal_loss_sum(m * k, xs.data_ptr(), ts.data_ptr(), *args, k, None,
                                                tot.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream) == 1
    assert abs(float(tot.item()) - float(l_rf.astype(np.float64).sum())) <= RTOL * abs(float(l_rf.sum())) + 1e-6


@pytest.mark.parametrize("m,k", [(500, 9), (30720, 2), (11, 81)])
def test_softmax_focal_parity(cuda_lib, ref_lib, oracle_mod, m, k):
    x, t = _inputs.focal_inputs(m, k, 4, softmax=True)
    args = (23.0, 2.0, 0.25)
    l, p, dx, buff = G.softmax_focal(cuda_lib, x, t, *args)
    l_rf, p_rf, dx_rf, buff_rf = G.softmax_focal(ref_lib, x, t, *args)
    np.testing.assert_allclose(p, p_rf, rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(l, l_rf, rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(buff, buff_rf, rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(dx, dx_rf, rtol=RTOL, atol=1e-9)
    lo, po = oracle_mod.softmax_focal_forward(x, t, *args)
    np.testing.assert_allclose(l, lo, rtol=RTOL, atol=1e-8)
    np.testing.assert_allclose(p, po, rtol=1e-5, atol=1e-9)


def test_autograd_wrappers(cuda_lib, oracle_mod):
    """The reference-facing Python classes: forward values and gradients."""
    import torch
    from scda_b200.extensions import RoIPool
    from scda_b200.extensions._roi_align.modules.roi_align import RoIAlign, RoIAlignAvg, RoIAlignMax
    from scda_b200.extensions._focal_loss.focal_loss import (SigmoidFocalLossFunction,
                                                             SoftmaxFocalLossFunction)
    feat_np = _inputs.features((1, 16, 32, 64), 0)
    rois_np = _inputs.rois_uniform(20, 1, img_w=1024, img_h=512)
    feat = torch.from_numpy(feat_np).cuda().requires_grad_(True)
    rois = torch.from_numpy(rois_np).cuda()
    out = RoIPool(7, 7, 1 / 16.)(feat, rois)
    g = torch.from_numpy(_inputs.features(tuple(out.shape), 2)).cuda()
    out.backward(g)
    o_or, a_or = oracle_mod.roi_pool_forward(feat_np, rois_np, 7, 7, 1 / 16.)
    assert np.array_equal(out.detach().cpu().numpy(), o_or)
    gi_or = oracle_mod.roi_pool_backward(g.cpu().numpy(), rois_np, a_or, feat_np.shape, 1 / 16.)
    np.testing.assert_allclose(feat.grad.cpu().numpy(), gi_or, rtol=RTOL, atol=1e-5)

    feat.grad = None
    out = RoIAlign(7, 7, 1 / 16.)(feat, rois)
    out.backward(g)
    np.testing.assert_allclose(out.detach().cpu().numpy(),
                               oracle_mod.roi_align_forward(feat_np, rois_np, 7, 7, 1 / 16.),
                               rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(feat.grad.cpu().numpy(),
                               oracle_mod.roi_align_backward(g.cpu().numpy(), rois_np, feat_np.shape, 1 / 16.),
                               rtol=RTOL, atol=1e-4)
    a8 = oracle_mod.roi_align_forward(feat_np, rois_np, 8, 8, 1 / 16.)
    avg = RoIAlignAvg(7, 7, 1 / 16.)(feat, rois).detach().cpu().numpy()
    mx = RoIAlignMax(7, 7, 1 / 16.)(feat, rois).detach().cpu().numpy()
    ref_avg = (a8[:, :, :-1, :-1] + a8[:, :, :-1, 1:] + a8[:, :, 1:, :-1] + a8[:, :, 1:, 1:]) / 4
    ref_max = np.maximum(np.maximum(a8[:, :, :-1, :-1], a8[:, :, :-1, 1:]),
                         np.maximum(a8[:, :, 1:, :-1], a8[:, :, 1:, 1:]))
    np.testing.assert_allclose(avg, ref_avg, rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(mx, ref_max, rtol=RTOL, atol=ATOL)

    x_np, t_np = _inputs.focal_inputs(400, 8, 5)
    x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    t = torch.from_numpy(t_np).cuda()
    wp = torch.tensor([19.0])
    loss = SigmoidFocalLossFunction(2.0, 0.25, 8)(x, t, wp)
    assert loss.shape == (1,) and loss.is_cuda
    (loss * 0.5).backward()
    ref_l = oracle_mod.sigmoid_focal_forward(x_np, t_np, 19.0, 2.0, 0.25).astype(np.float64).sum()
    assert abs(float(loss.item()) - ref_l) <= RTOL * abs(ref_l)
    np.testing.assert_allclose(x.grad.cpu().numpy(),
                               0.5 * oracle_mod.sigmoid_focal_backward(x_np, t_np, 19.0, 2.0, 0.25),
                               rtol=RTOL, atol=1e-8)
    x_np, t_np = _inputs.focal_inputs(400, 9, 6, softmax=True)
    x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    t = torch.from_numpy(t_np).cuda()
    loss = SoftmaxFocalLossFunction(2.0, 0.25, 9)(x, t, wp)
    loss.backward()
    lo, po = oracle_mod.softmax_focal_forward(x_np, t_np, 19.0, 2.0, 0.25)
    assert abs(float(loss.item()) - lo.astype(np.float64).sum()) <= RTOL * abs(lo.sum())
    dxo, _ = oracle_mod.softmax_focal_backward(x_np, t_np, po, 19.0, 2.0, 0.25)
    np.testing.assert_allclose(x.grad.cpu().numpy(), dxo, rtol=RTOL, atol=1e-8)


def test_roi_pool_model_shape_properties(cuda_lib):
    """The model's operating point (1x512x32x64, 512 RoIs): argmax points at the
    value; backward conserves gradient mass and equals a dense scatter."""
    import torch
    feat = torch.from_numpy(_inputs.features((1, 512, 32, 64), 0)).cuda()
    rois = torch.from_numpy(_inputs.rois_uniform(512, 1, img_w=1024, img_h=512)).cuda()
    out = torch.empty(512, 512, 7, 7, device="cuda")
    arg = torch.empty(512, 512, 7, 7, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    assert cuda_lib.ROIPoolForwardLaucher(feat.data_ptr(), 1 / 16., 512, 32, 64, 512, 7, 7,
                                          rois.data_ptr(), out.data_ptr(), arg.data_ptr(), s) == 1
    ok = arg >= 0
    assert torch.equal(feat.flatten()[arg[ok].long()], out[ok])
    assert bool((out[~ok] == 0).all())
    g = torch.randn_like(out)
    gi = torch.empty_like(feat)
    assert cuda_lib.ROIPoolBackwardLaucher(g.data_ptr(), 1 / 16., 1, 512, 32, 64, 512, 7, 7,
                                           rois.data_ptr(), gi.data_ptr(), arg.data_ptr(), s) == 1
    dense = torch.zeros(feat.numel(), dtype=torch.float64, device="cuda")
    dense.index_add_(0, arg[ok].long(), g[ok].double())
    assert float((gi.flatten().double() - dense).abs().max()) < 1e-3
