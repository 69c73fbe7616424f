"""Device-side proposal / target plumbing against the oracle's numpy restatement of the
reference (oracle/host.py, itself pinned to the reference's own Python by
tests/test_oracle_host.py) on the same seeded inputs and the same sampling keys."""
import os

import numpy as np
import pytest

import _inputs
from _sampling_adapter import KeyedChoice

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_plumbing.npz")


@pytest.fixture(scope="module")
def cfg():
    return _inputs.load_cfg()


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def host():
    from oracle import host
    return host


def canon(rows):
    """Order rows inside groups of EQUAL score by their coordinates.  The reference orders
    ties by numpy's argpartition + unstable argsort (functions/rpn_proposal.py:53-55), i.e.
    arbitrarily; the device top-k orders them by index.  Everything else must agree."""
    rows = np.asarray(rows)
    key = np.lexsort((np.round(rows[:, 4], 2), np.round(rows[:, 3], 2), np.round(rows[:, 2], 2),
                      np.round(rows[:, 1], 2), -rows[:, 5].astype(np.float64)))
    return rows[key]


def test_anchors_device(cuda_lib, g, cfg):
    from scda_b200.utils import anchor_helper
    sh = cfg["shared"]
    a = anchor_helper.anchors_device(32, 64, sh["anchor_ratios"], sh["anchor_scales"], 16, "cuda")
    assert np.array_equal(a.cpu().numpy(), g["anchors_32x64"])


@pytest.mark.parametrize("tag", ["train", "test"])
def test_rpn_proposals_match_reference_golden(cuda_lib, g, cfg, tag):
    import torch
    from scda_b200.functions.rpn_proposal import compute_rpn_proposals
    cls, loc = _inputs.synth_rpn_outputs(0)
    out = compute_rpn_proposals(torch.from_numpy(cls).cuda(), torch.from_numpy(loc).cuda(),
                                cfg[tag + "_rpn_proposal_cfg"], torch.from_numpy(g["image_info"]))
    ref = g["proposals_" + tag]
    assert out.device.type == "cpu" and out.dtype == torch.float32
    assert out.shape == ref.shape
    o = out.numpy()
    # scores and the kept set are exact; box corners go through exp() (libdevice vs numpy:
    # <= 1 ulp in float32 before the float64 products), so allow 1e-5 relative there
    assert np.array_equal(o[:, 0], ref[:, 0]) and np.array_equal(o[:, 5], ref[:, 5])
    o, ref = canon(o), canon(ref)
    np.testing.assert_allclose(o[:, 1:5], ref[:, 1:5], rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_rpn_proposals_vs_oracle_other_seeds(cuda_lib, cfg, host, seed):
    import torch
    from scda_b200.functions.rpn_proposal import compute_rpn_proposals
    cls, loc = _inputs.synth_rpn_outputs(10 + seed, fh=20, fw=30)
    info = np.array([[320, 480, 1.0]], np.float32)
    c = dict(cfg["train_rpn_proposal_cfg"], pre_nms_top_n=3000, post_nms_top_n=500, roi_min_size=12)
    ref = host.compute_rpn_proposals(cls, loc, c, info)
    out = compute_rpn_proposals(torch.from_numpy(cls).cuda(), torch.from_numpy(loc).cuda(), c,
                                torch.from_numpy(info)).numpy()
    assert out.shape == ref.shape
    assert np.array_equal(out[:, 5], ref[:, 5])
    out, ref = canon(out), canon(ref)
    np.testing.assert_allclose(out[:, 1:5], ref[:, 1:5], rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("KA,pre,ties,min_size", [
    (30720, 12000, False, 16.0),      # the training configuration
    (30720, 6000, True, 16.0),        # quantised scores: long runs of equal keys across the selection boundary
    (9000, 0, True, 2.0),             # pre <= 0: every anchor, sorted
    (5001, 5001, False, 40.0),        # K not a multiple of four, many rows below the minimum size
    (700, 33, True, 0.0),
])
def test_rpn_proposal_rows_kernel(cuda_lib, KA, pre, ties, min_size):
    """scda_rpn_proposal_rows (selection + order by rank counting + decode + compaction, two launches) against
    a stable descending sort on the host followed by the single-CTA decode kernel (scda_rpn_decode_pack, itself
    pinned to the reference's golden proposals above): identical rows, bit for bit, ties by anchor index."""
    import torch
    from scda_b200._lib import check, load, stream_ptr
    lib = load()
    r = np.random.RandomState(KA + pre)
    scores = r.uniform(0, 1, KA).astype(np.float32)
    if ties:
        scores = np.round(scores * 50) / 50
        scores[r.randint(0, KA, KA // 10)] *= -1            # negative keys too
        scores = scores + np.float32(0)                     # no -0.0: numpy ties it with +0.0
    anchors = np.sort(r.uniform(0, 900, (KA, 2, 2)), axis=1).reshape(KA, 4)[:, [0, 2, 1, 3]].astype(np.float64)
    deltas = (r.standard_normal((KA, 4)) * 0.3).astype(np.float32)
    K = KA if pre <= 0 or pre > KA else pre
    order = np.argsort(-scores.astype(np.float64), kind="stable")[:K]
    dev = "cuda"
    t_scores, t_anchors, t_deltas = (torch.from_numpy(x).to(dev) for x in (scores, anchors, deltas))
    t_order = torch.from_numpy(order.astype(np.int64)).to(dev)
    t_top = torch.from_numpy(scores[order]).to(dev)
    want = torch.full((K, 5), 7.0, device=dev)
    want_n = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib.scda_rpn_decode_pack(K, t_anchors.data_ptr(), t_deltas.data_ptr(), t_order.data_ptr(), t_top.data_ptr(),
                                   600.0, 1000.0, min_size, want.data_ptr(), want_n.data_ptr(), stream_ptr(dev)), "ref")
    got = torch.full((K, 5), 9.0, device=dev)
    got_n = torch.zeros(1, dtype=torch.int32, device=dev)
    wsb = lib.scda_rpn_proposal_rows_workspace_bytes(KA, pre)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    for _ in range(2):                                       # the ticket re-arms itself
        got.fill_(9.0)
        check(lib.scda_rpn_proposal_rows(KA, pre, t_scores.data_ptr(), t_anchors.data_ptr(), t_deltas.data_ptr(),
                                         600.0, 1000.0, min_size, got.data_ptr(), got_n.data_ptr(), ws.data_ptr(),
                                         wsb, stream_ptr(dev)), "scda_rpn_proposal_rows")
        torch.cuda.synchronize()
        assert int(got_n) == int(want_n)
        assert 0 < int(got_n) <= K
        assert torch.equal(got, want)
    if min_size >= 40:
        assert int(got_n) < K


def _keys(seed, n, tags):
    r = np.random.RandomState(seed)
    return {t: r.uniform(0, 1, n) for t in tags}


@pytest.mark.parametrize("seed,G", [(0, 20), (1, 3), (2, 60)])
def test_anchor_targets_vs_oracle(cuda_lib, cfg, host, seed, G):
    import torch
    from scda_b200.functions.anchor_target import compute_anchor_targets
    from scda_b200.functions._sampling import ArrayRng
    gts = _inputs.gt_boxes(G, seed)[None]
    gts = np.concatenate([gts, np.zeros((1, 4, 5), np.float32)], 1)      # padded rows, like the loader
    info = np.array([[512, 1024, 0.5]], np.float32)
    c = cfg["train_anchor_target_cfg"]
    keys = _keys(seed, 30720, ["pos", "neg"])
    ref = host.compute_anchor_targets((1, 60, 32, 64), c, gts, info, choice=KeyedChoice(keys))
    out = compute_anchor_targets((1, 60, 32, 64), c, torch.from_numpy(gts).cuda(),
                                 torch.from_numpy(info), rng=ArrayRng([keys["pos"], keys["neg"]]))
    ct, lt, lm, norm = [o.cpu().numpy() for o in out]
    assert np.array_equal(ct, ref[0])
    assert np.array_equal(lm, ref[2])
    np.testing.assert_allclose(lt, ref[1], rtol=1e-6, atol=1e-7)
    assert int(norm) == ref[3]
    assert (ct == 1).sum() <= 128 and (ct >= 0).sum() <= 256


def test_anchor_targets_match_reference_golden_sets(cuda_lib, cfg, g):
    """Against the reference's own run (np.random.seed(123)): the sampled subsets differ (other
    RNG), the pre-sampling label sets and the loc targets of common positives must not."""
    import torch
    from scda_b200.functions.anchor_target import compute_anchor_targets
    c = dict(cfg["train_anchor_target_cfg"], rpn_batch_size=10 ** 6)     # no sub-sampling
    out = compute_anchor_targets((1, 60, 32, 64), c, torch.from_numpy(g["gts"]).cuda(),
                                 torch.from_numpy(g["image_info"]))
    ct, lt = out[0].cpu().numpy(), out[1].cpu().numpy()
    ref_ct, ref_lt = g["anchor_cls_targets"], g["anchor_loc_targets"]
    assert np.all(ct[ref_ct == 1] == 1) and np.all(ct[ref_ct == 0] == 0)   # sampled ⊂ full sets
    pos = np.repeat(ref_ct == 1, 4, axis=1)
    np.testing.assert_allclose(lt[pos], ref_lt[pos], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("seed,npos_jitter", [(0, 30), (1, 2), (2, 0)])
def test_proposal_targets_vs_oracle(cuda_lib, cfg, host, g, seed, npos_jitter):
    import torch
    from scda_b200.functions.proposal_target import compute_proposal_targets
    from scda_b200.functions._sampling import ArrayRng
    r = np.random.RandomState(seed)
    gts = g["gts"]
    base = g["proposals_train"][:900 if seed != 2 else 200, 1:5]
    jit = np.repeat(gts[0, :, :4], npos_jitter, axis=0) + r.normal(0, 6, (20 * npos_jitter, 4)).astype(np.float32)
    props = np.vstack([base, jit]).astype(np.float32)
    props = np.hstack([np.zeros((len(props), 1), np.float32), props, np.zeros((len(props), 1), np.float32)])
    info = g["image_info"]
    c = cfg["train_proposal_target_cfg"]
    keys = _keys(100 + seed, 4096, ["pos", "neg", "pad"])
    ref = host.compute_proposal_targets(props.copy(), c, gts, info, choice=KeyedChoice(keys))
    out = compute_proposal_targets(torch.from_numpy(props).cuda(), c, torch.from_numpy(gts).cuda(),
                                   torch.from_numpy(info),
                                   rng=ArrayRng([keys["pos"], keys["neg"], keys["pad"]]))
    rois, lab, t, w = [o.cpu().numpy() for o in out]
    assert rois.shape == (512, 5) and out[0].is_cuda
    assert np.array_equal(rois, ref[0])
    assert np.array_equal(lab, ref[1])
    assert np.array_equal(w, ref[3])
    np.testing.assert_allclose(t, ref[2], rtol=1e-5, atol=1e-6)


def test_predicted_bboxes_match_reference_golden(cuda_lib, cfg, g):
    import torch
    from scda_b200.functions.predict_bbox import compute_predicted_bboxes
    rois = torch.from_numpy(g["proposals_test"][:, :5].copy()).cuda()
    out = compute_predicted_bboxes(rois, torch.from_numpy(g["pb_cls"]).cuda(),
                                   torch.from_numpy(g["pb_loc"]).cuda(), g["image_info"],
                                   cfg["test_predict_bbox_cfg"]).cpu().numpy()
    ref = g["pb_out"]
    assert out.shape == ref.shape
    assert np.array_equal(out[:, 5], ref[:, 5])              # scores, in order
    key = lambda r: r[np.lexsort((r[:, 6], -r[:, 5].astype(np.float64)))]
    out, ref = key(out), key(ref)
    assert np.array_equal(out[:, 6], ref[:, 6])              # classes
    np.testing.assert_allclose(out[:, 1:5], ref[:, 1:5], rtol=1e-5, atol=1e-4)


def test_cluster_targets(cuda_lib, host):
    import torch
    from scda_b200.functions.mask import compute_cluster_targets
    rois = _inputs.rois_uniform(512, 3, img_w=1024, img_h=512, wh=(16, 200))
    fea = np.random.RandomState(0).standard_normal((512, 64)).astype(np.float32)
    ref, ref_c, labels = host.compute_cluster_targets(rois, fea, 4, 128)
    out, centers = compute_cluster_targets(torch.from_numpy(rois).cuda(), torch.from_numpy(fea).cuda().requires_grad_(True), 4, 128)
    assert out.shape == (4, 128, 64) and not out.requires_grad
    np.testing.assert_allclose(centers, ref_c, rtol=1e-5, atol=1e-3)
    out = out.cpu().numpy()
    for c in range(4):
        members = np.where(labels == c)[0]
        if members.shape[0] >= 128:
            assert np.array_equal(out[c], ref[c])            # first 128 members, in RoI order
        else:                                                 # re-drawn with replacement: any member
            rows = {fea[i].tobytes() for i in members}
            assert all(r.tobytes() in rows for r in out[c])
