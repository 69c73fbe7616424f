"""CPU suite: the oracle against the committed golden vectors (outputs of the
reference's own code, tests/golden/make_golden.py) and against independent
numpy restatements / properties."""
import glob
import importlib.util
import os

import numpy as np
import pytest

import _inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gold(name):
    p = os.path.join(GOLD, name + ".npz")
    if not os.path.exists(p):
        pytest.skip("golden %s not generated yet" % name)
    return np.load(p)


def upper_words(mask):
    n, cb = mask.shape
    rb = (np.arange(n) // 64)[:, None]
    return np.where(np.arange(cb)[None, :] >= rb, mask, 0)


# ------------------------------------------------------------- golden vectors
def test_bbox_overlaps_matches_reference_cython_golden(oracle_mod):
    g = gold("cython_bbox_overlaps")
    assert np.array_equal(oracle_mod.bbox_overlaps(g["boxes"], g["query"]), g["overlaps"])


def test_bbox_overlaps_matches_live_reference_cython(oracle_mod):
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "cython_bbox*.so"))
    if not so:
        pytest.skip("reference cython_bbox not built")
    spec = importlib.util.spec_from_file_location("cython_bbox", so[0])
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    a = _inputs.nms_boxes(4000, 3)[:, :4].copy()
    b = _inputs.gt_boxes(50, 4)[:, :4].copy()
    assert np.array_equal(oracle_mod.bbox_overlaps(a, b), m.bbox_overlaps(a, b))


def test_roi_pool_matches_reference_kernel_golden(oracle_mod):
    g = gold("roi_pool")
    out, arg = oracle_mod.roi_pool_forward(g["feat"], g["rois"], 7, 7, float(g["scale"]))
    assert np.array_equal(out, g["out"])
    assert np.array_equal(arg, g["argmax"])
    gi = oracle_mod.roi_pool_backward(g["top_diff"], g["rois"], g["argmax"], g["feat"].shape,
                                      float(g["scale"]))
    assert np.array_equal(gi, g["bottom_diff"])      # same summation order as the gather kernel


def test_roi_align_matches_reference_kernel_golden(oracle_mod):
    g = gold("roi_align")
    out = oracle_mod.roi_align_forward(g["feat"], g["rois"], 8, 8, float(g["scale"]))
    np.testing.assert_allclose(out, g["out"], rtol=1e-6, atol=1e-7)
    gi = oracle_mod.roi_align_backward(g["top_diff"], g["rois"], g["feat"].shape, float(g["scale"]))
    np.testing.assert_allclose(gi, g["bottom_diff"], rtol=1e-4, atol=1e-5)   # atomic order differs


@pytest.mark.parametrize("name", ["nms_uniform", "nms_clustered"])
def test_nms_matches_reference_kernel_golden(oracle_mod, name):
    g = gold(name)
    mask = oracle_mod.nms_mask(g["boxes"], float(g["thresh"]))
    assert np.array_equal(upper_words(mask), upper_words(g["mask"]))
    assert np.array_equal(oracle_mod.nms(g["boxes"], float(g["thresh"])), g["keep"])


def test_iou_overlap_matches_reference_kernel_golden(oracle_mod):
    g = gold("iou_overlap")
    assert np.array_equal(oracle_mod.iou_overlap(g["b1"], g["b2"]), g["out"])


def test_focal_matches_reference_kernel_golden(oracle_mod):
    g = gold("sigmoid_focal")
    a = (float(g["weight_pos"]), float(g["gamma"]), float(g["alpha"]))
    np.testing.assert_allclose(oracle_mod.sigmoid_focal_forward(g["logits"], g["targets"], *a),
                               g["losses"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(oracle_mod.sigmoid_focal_backward(g["logits"], g["targets"], *a),
                               g["dx"], rtol=1e-4, atol=1e-7)
    g = gold("softmax_focal")
    l, p = oracle_mod.softmax_focal_forward(g["logits"], g["targets"], *a)
    np.testing.assert_allclose(l, g["losses"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(p, g["priors"], rtol=1e-5, atol=1e-8)
    dx, buff = oracle_mod.softmax_focal_backward(g["logits"], g["targets"], g["priors"], *a)
    np.testing.assert_allclose(dx, g["dx"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(buff, g["buff"], rtol=1e-4, atol=1e-7)


# -------------------------------------------- independent restatements / props
def _py_roi_pool(feat, rois, ph, pw, scale):
    """Plain-Python loop of roi_pooling_kernel.cu:39-91 (small cases only)."""
    f32 = np.float32
    R, (B, C, H, W) = rois.shape[0], feat.shape
    out = np.zeros((R, C, ph, pw), f32)
    arg = np.full((R, C, ph, pw), -1, np.int32)

    def rnd(v):  # round half away from zero, as roundf
        return int(np.floor(abs(v) + f32(0.5)) * (1 if v >= 0 else -1))
    for n in range(R):
        b = int(rois[n, 0])
        x0, y0, x1, y1 = [rnd(f32(rois[n, k]) * f32(scale)) for k in (1, 2, 3, 4)]
        rw, rh = max(x1 - x0 + 1, 1), max(y1 - y0 + 1, 1)
        bh, bw = f32(rh) / f32(ph), f32(rw) / f32(pw)
        for i in range(ph):
            hs = min(max(int(np.floor(f32(i) * bh)) + y0, 0), H)
            he = min(max(int(np.ceil(f32(i + 1) * bh)) + y0, 0), H)
            for j in range(pw):
                ws = min(max(int(np.floor(f32(j) * bw)) + x0, 0), W)
                we = min(max(int(np.ceil(f32(j + 1) * bw)) + x0, 0), W)
                if he <= hs or we <= ws:
                    out[n, :, i, j] = 0
                    continue
                win = feat[b, :, hs:he, ws:we].reshape(C, -1)
                k = win.argmax(1)
                out[n, :, i, j] = win[np.arange(C), k]
                hh, ww = hs + k // (we - ws), ws + k % (we - ws)
                arg[n, :, i, j] = ((b * C + np.arange(C)) * H + hh) * W + ww
    return out, arg


def test_roi_pool_forward_vs_python_restatement(oracle_mod):
    feat = _inputs.features((2, 5, 12, 17), 0)
    rois = _inputs.rois_uniform(40, 1, img_w=17 * 16, img_h=12 * 16, wh=(4, 200), batch=2)
    rois[0, 1:] = [-30, -30, 20, 20]
    rois[1, 1:] = [90, 90, 80, 80]
    out, arg = oracle_mod.roi_pool_forward(feat, rois, 7, 7, 1 / 16.)
    o2, a2 = _py_roi_pool(feat, rois, 7, 7, 1 / 16.)
    assert np.array_equal(out, o2)
    assert np.array_equal(arg, a2)


def test_roi_pool_properties(oracle_mod):
    feat = _inputs.features((1, 8, 32, 64), 0)
    rois = _inputs.rois_uniform(64, 1, img_w=1024, img_h=512)
    out, arg = oracle_mod.roi_pool_forward(feat, rois, 7, 7, 1 / 16.)
    ok = arg >= 0
    assert np.array_equal(feat.ravel()[arg[ok]], out[ok])          # argmax points at the value
    assert np.all(out[~ok] == 0)
    g = _inputs.features(out.shape, 2)
    gi = oracle_mod.roi_pool_backward(g, rois, arg, feat.shape, 1 / 16.)
    ref = np.zeros(feat.size, np.float64)
    np.add.at(ref, arg[ok], g[ok].astype(np.float64))
    np.testing.assert_allclose(gi.ravel(), ref, rtol=1e-5, atol=1e-5)
    assert abs(gi.sum() - g[ok].sum()) < 1e-2                      # gradient mass is conserved
    # no-argmax form produces the same output
    assert np.array_equal(oracle_mod.roi_pool_forward(feat, rois, 7, 7, 1 / 16., want_argmax=False), out)


def test_roi_pool_empty_and_single(oracle_mod):
    feat = _inputs.features((1, 3, 8, 8), 0)
    out, arg = oracle_mod.roi_pool_forward(feat, np.zeros((0, 5), np.float32), 7, 7, 1 / 16.)
    assert out.shape == (0, 3, 7, 7)
    # an RoI covering the whole map with a 1x1 pool is the channel max
    roi = np.array([[0, 0, 0, 127, 127]], np.float32)
    out, _ = oracle_mod.roi_pool_forward(feat, roi, 1, 1, 1 / 16.)
    assert np.array_equal(out[0, :, 0, 0], feat[0].reshape(3, -1).max(1))


def _np_roi_align(feat, rois, ah, aw, scale):
    """float64 restatement of the sampling rule (roi_align_kernel.cu:33-68)."""
    R, (B, C, H, W) = rois.shape[0], feat.shape
    out = np.zeros((R, C, ah, aw))
    for n in range(R):
        b = int(rois[n, 0])
        x0, y0, x1, y1 = [float(np.float32(rois[n, k]) * np.float32(scale)) for k in (1, 2, 3, 4)]
        rw, rh = max(x1 - x0 + 1, 0), max(y1 - y0 + 1, 0)
        for i in range(ah):
            for j in range(aw):
                h, w = i * rh / (ah - 1) + y0, j * rw / (aw - 1) + x0
                if h < 0 or h >= H or w < 0 or w >= W:
                    continue
                hs, ws = int(min(np.floor(h), H - 2)), int(min(np.floor(w), W - 2))
                hr, wr = h - hs, w - ws
                f = feat[b].astype(np.float64)
                out[n, :, i, j] = (f[:, hs, ws] * (1 - hr) * (1 - wr) + f[:, hs, ws + 1] * (1 - hr) * wr
                                   + f[:, hs + 1, ws] * hr * (1 - wr) + f[:, hs + 1, ws + 1] * hr * wr)
    return out


def test_roi_align_config1_vs_numpy(oracle_mod):
    """BASELINE.json configs[0]: 1x256x64x64 map, 128 RoIs (channel count cut for the
    Python loop; the C oracle runs the full 256)."""
    feat = _inputs.features((1, 256, 64, 64), 0)
    rois = _inputs.rois_uniform(128, 1)
    out = oracle_mod.roi_align_forward(feat, rois, 7, 7, 1 / 16.)
    ref = _np_roi_align(feat[:, :16], rois, 7, 7, 1 / 16.)
    np.testing.assert_allclose(out[:, :16], ref, rtol=1e-4, atol=1e-5)
    # backward is the adjoint of forward: <fwd(x), g> == <x, bwd(g)>
    g = _inputs.features(out.shape, 1)
    gi = oracle_mod.roi_align_backward(g, rois, feat.shape, 1 / 16.)
    lhs = float((out.astype(np.float64) * g).sum())
    rhs = float((feat.astype(np.float64) * gi).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0)


def _np_greedy_nms(boxes, thresh, ge=False):
    x1, y1, x2, y2 = [boxes[:, k].astype(np.float32) for k in range(4)]
    n = len(boxes)
    dead = np.zeros(n, bool)
    keep = []
    area = (x2 - x1 + 1) * (y2 - y1 + 1)
    for i in range(n):
        if dead[i]:
            continue
        keep.append(i)
        w = np.maximum(np.minimum(x2[i], x2) - np.maximum(x1[i], x1) + 1, 0)
        h = np.maximum(np.minimum(y2[i], y2) - np.maximum(y1[i], y1) + 1, 0)
        iou = w * h / (area[i] + area - w * h)
        sup = (iou >= thresh) if ge else (iou > thresh)
        sup[: i + 1] = False
        dead |= sup
    return np.array(keep, np.int64)


@pytest.mark.parametrize("n,seed,th", [(1, 0, 0.7), (63, 1, 0.7), (64, 2, 0.5), (65, 3, 0.7),
                                       (1000, 4, 0.7), (3000, 5, 0.5)])
def test_nms_mask_scan_equals_greedy(oracle_mod, n, seed, th):
    b = _inputs.nms_boxes(n, seed) if seed % 2 == 0 else _inputs.clustered_boxes(n, seed)
    k1 = oracle_mod.nms_scan(oracle_mod.nms_mask(b, th))
    k2 = oracle_mod.nms(b, th)
    assert np.array_equal(k1, k2)
    # the numpy version rounds Sa+Sb without the fma; decisions may flip only at ties
    k3 = _np_greedy_nms(b, th)
    assert len(set(k1) ^ set(k3)) <= max(2, n // 200)
    # idempotent: NMS of the survivors keeps all of them
    assert np.array_equal(oracle_mod.nms(b[k1], th), np.arange(len(k1)))
    assert np.all(np.diff(k1) > 0)


def test_nms_empty_and_identical(oracle_mod):
    assert oracle_mod.nms(np.zeros((0, 5), np.float32), 0.7).size == 0
    b = np.tile(np.array([[10, 10, 50, 50, 0.9]], np.float32), (130, 1))
    assert np.array_equal(oracle_mod.nms(b, 0.7), [0])
    # thresh 1.0 with strict '>' suppresses nothing, even exact duplicates
    assert len(oracle_mod.nms(b, 1.0)) == 130
    # cpu_nms semantics ('>=') do suppress at IoU == thresh
    assert np.array_equal(oracle_mod.cpu_nms(b, 1.0), [0])


def test_iou_conventions_differ_as_documented(oracle_mod):
    a = np.array([[0, 0, 10, 10], [0, 0, 0.5, 0.5]], np.float32)
    q = np.array([[5, 5, 15, 15], [10, 0, 20, 10], [0, 0, 0.5, 0.5]], np.float32)
    cy = oracle_mod.bbox_overlaps(a, q)
    gp = oracle_mod.iou_overlap(a, q)
    assert cy[0, 0] == np.float32(25) / np.float32(175) and gp[0, 0] == cy[0, 0]
    assert cy[0, 1] == 0 and gp[0, 1] == 0            # touching edges: zero overlap
    assert cy[1, 2] == 1.0                            # tiny identical boxes: IoU 1 ...
    assert gp[1, 2] == np.float32(0.25)               # ... but the GPU helper clamps union to 1
    assert np.array_equal(oracle_mod.bbox_overlaps(q, q), oracle_mod.bbox_overlaps(q, q).T)


def test_sigmoid_focal_vs_float64_formula(oracle_mod):
    x, t = _inputs.focal_inputs(300, 8, 0)
    wp, gamma, alpha = 17.0, 2.0, 0.25
    l = oracle_mod.sigmoid_focal_forward(x, t, wp, gamma, alpha)
    xd = x.astype(np.float64)
    p = 1 / (1 + np.exp(-xd))
    d = np.arange(8)[None, :]
    c1 = (t[:, None] == d + 1)
    c2 = (t[:, None] != -1) & ~c1
    ref = -(c1 * (1 - p) ** gamma * np.log(p) * alpha / wp) - (c2 * p ** gamma * np.log1p(-p) * (1 - alpha) / wp)
    np.testing.assert_allclose(l, ref, rtol=2e-4, atol=1e-7)
    # gradient == finite differences of the float64 formula
    dx = oracle_mod.sigmoid_focal_backward(x, t, wp, gamma, alpha)
    eps = 1e-6

    def f(z):
        pz = 1 / (1 + np.exp(-z))
        return -(c1 * (1 - pz) ** gamma * np.log(pz) * alpha / wp) - (c2 * pz ** gamma * np.log1p(-pz) * (1 - alpha) / wp)
    fd = (f(xd + eps) - f(xd - eps)) / (2 * eps)
    np.testing.assert_allclose(dx, fd, rtol=1e-3, atol=1e-6)


def test_softmax_focal_vs_float64_formula(oracle_mod):
    x, t = _inputs.focal_inputs(300, 9, 1, softmax=True)
    wp, gamma, alpha = 17.0, 2.0, 0.25
    l, P = oracle_mod.softmax_focal_forward(x, t, wp, gamma, alpha)
    xd = x.astype(np.float64)
    e = np.exp(xd - xd.max(1, keepdims=True))
    Pd = e / e.sum(1, keepdims=True)
    np.testing.assert_allclose(P, Pd, rtol=1e-5, atol=1e-8)
    idx = np.where(t >= 0)[0]
    pl = Pd[idx, t[idx]]
    z = np.where(t[idx] == 0, (1 - alpha) / wp, alpha / wp)
    ref = np.zeros(len(t))
    ref[idx] = -((1 - pl) ** gamma) * np.log(pl) * z
    np.testing.assert_allclose(l, ref, rtol=2e-4, atol=1e-7)
    dx, _ = oracle_mod.softmax_focal_backward(x, t, P, wp, gamma, alpha)
    assert np.all(dx[t < 0] == 0)
    np.testing.assert_allclose(dx.sum(1), 0, atol=1e-6)           # softmax gradients sum to zero
