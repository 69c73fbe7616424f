"""One SCDA training iteration (scda_b200/engine.py) on the GPU: it runs and updates all four
networks; the CUDA-graph replay of the reconstruction / discriminator phases equals the
eager execution; the weight gradients the tensor-core backward writes straight into the
flat gradient buffer equal the ones it hands to autograd."""
import numpy as np
import pytest

import _inputs

pytestmark = pytest.mark.gpu

H, W = 256, 512


def _batch(seed=0):
    import torch
    r = np.random.RandomState(seed)
    img = torch.from_numpy(r.standard_normal((1, 3, H, W)).astype(np.float32)).cuda()
    tgt = torch.from_numpy(r.standard_normal((1, 3, H, W)).astype(np.float32)).cuda()
    gts = torch.from_numpy(_inputs.gt_boxes(12, seed, img_w=W, img_h=H)[None]).cuda()
    info = torch.tensor([[H, W, 0.5]])
    return img, tgt, gts, info


def _trainer(use_graphs, lr=1e-4, deterministic=False):
    import torch
    from scda_b200.engine import build_trainer
    cfg = _inputs.load_cfg()
    tr = build_trainer(cfg, lr=lr, new_w=W, new_h=H, world_size=1, seed=0, use_graphs=use_graphs)
    if deterministic:
        for net in tr.nets():
            for m in net.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
    return tr, cfg


def test_iteration_updates_all_four_networks(cuda_lib):
    import torch
    tr, cfg = _trainer(True)
    img, tgt, gts, info = _batch()
    before = [net.state_dict()[next(iter(net.state_dict()))].clone() for net in tr.nets()]
    outs = []
    for it in range(3):           # 1st eager, 2nd captures + replays, 3rd replays
        torch.manual_seed(it)
        np.random.seed(it)
        outs.append(tr.iteration(cfg, img, info, gts, tgt))
    torch.cuda.synchronize()
    for o in outs:
        for k, v in o.items():
            assert bool(torch.isfinite(v).all()), k
    after = [net.state_dict()[next(iter(net.state_dict()))] for net in tr.nets()]
    for b, a in zip(before, after):
        assert not torch.equal(a, b), "a network was not updated"
    assert tr.opt.t == 3 and tr.opt_dec.t == 3 and tr.opt_dis.t == 3 and tr.opt_dis_patch.t == 3
    assert tr._graphs is not None
    # bf16 shadows follow the fp32 masters
    w = tr.model.features[2].weight
    assert torch.equal(w._scda_shadow, w.detach().permute(0, 2, 3, 1).bfloat16())


def test_graph_replay_equals_eager(cuda_lib, monkeypatch):
    """dropout off, fixed soft labels and sampling keys, a learning rate small enough that the
    weights stay put: the captured (two-stream) iteration and the eager single-stream iteration
    compute the same losses and the same gradients on every one of three iterations.
    (Comparing Adam UPDATES at a real learning rate is ill-conditioned: Adam moves each weight by
    ~lr whatever the size of its gradient, so gradients at the fp32 / TF32 noise floor — e.g.
    the biases in front of an InstanceNorm, whose true gradient is zero — flip sign between two
    runs of cuDNN and the difference compounds.)"""
    import torch
    from scda_b200 import engine
    monkeypatch.setattr(engine, "soft_label",
                        lambda flag, like, generator=None: torch.full_like(like, 0.9 if flag == 1 else 0.15))
    # the Philox offsets a captured graph consumes differ from the eager ones for the same
    # seed, so every random draw (sampling keys, member re-draws) is made a constant here
    real_full = torch.full

    def fake_rand(*size, **kw):
        kw.pop("generator", None)
        size = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else size
        return real_full(tuple(size), 0.37, **kw)
    monkeypatch.setattr(torch, "rand", fake_rand)
    img, tgt, gts, info = _batch(1)
    results = []
    for use_graphs, overlap, cut in ((False, False, False), (True, True, False), (True, True, True)):
        from scda_b200.engine import build_trainer
        cfg = _inputs.load_cfg()
        tr = build_trainer(cfg, lr=1e-9, new_w=W, new_h=H, world_size=1, seed=0, use_graphs=use_graphs,
                           overlap=overlap, force_cut=cut)
        for net in tr.nets():
            for m in net.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
        losses, grads = [], []
        for it in range(3):
            torch.manual_seed(100 + it)
            np.random.seed(100 + it)
            out = tr.iteration(cfg, img, info, gts, tgt)
            losses.append({k: float(v) for k, v in out.items()})
            grads.append([o.bucket.flat.clone() for o in (tr.opt, tr.opt_dec, tr.opt_dis, tr.opt_dis_patch)])
        torch.cuda.synchronize()
        if use_graphs:
            # one graph, or (the plan used when world > 1) one stretch per segment on four streams
            assert tr._graphs is not None and len(tr._graphs) == (len(tr._segments()) if cut else 1)
        results.append((losses, grads))
    (l_e, g_e) = results[0]
    for mode, (l_g, g_g) in zip(("one graph", "cut"), results[1:]):
        for a, b in zip(l_e, l_g):
            for k in ("loss", "rpn_cls", "rpn_loc", "rcnn_cls", "rcnn_loc", "dis_loss", "dis_patch_loss",
                      "dec_loss", "fake_loss"):
                assert abs(a[k] - b[k]) <= 2e-3 * max(1.0, abs(a[k])), (mode, k, a[k], b[k])
        # gradients: bf16 tensor-core detector (deterministic kernels, atomics only in the RoI /
        # bias sums), TF32 cuDNN reconstruction networks (the algorithm may differ under capture)
        for it in range(3):
            for name, a, b, lim in zip(("detector", "decoder", "dis", "dis_patch"), g_e[it], g_g[it],
                                       (0.9999, 0.999, 0.999, 0.999)):
                cos = float((a * b).sum() / (a.norm() * b.norm()))
                assert cos > lim, (mode, it, name, cos)
                assert 0.99 < float(a.norm() / b.norm()) < 1.01, (mode, it, name, float(a.norm() / b.norm()))


def test_direct_gradient_sink_equals_autograd_gradients(cuda_lib):
    import torch
    tr, cfg = _trainer(False, deterministic=True)
    img, tgt, gts, info = _batch(2)
    x = {'cfg': cfg, 'image': img, 'image_info': info, 'ground_truth_bboxes': gts,
         'ignore_regions': None, 'cluster_num': 4, 'threshold': 128}

    def run():
        torch.manual_seed(7)
        np.random.seed(7)
        out = tr.model(x, tgt)
        return sum(out['losses'])

    tr.opt.zero_grad()
    run().backward(inputs=tr.opt.params)
    tr.opt.bucket.settle()
    direct = [p.grad.detach().clone() for p in tr.opt.params]
    assert float(direct[0].abs().sum()) > 0 and float(direct[-1].abs().sum()) > 0
    for p in tr.opt.params:
        p._scda_direct_grad = False
        p.grad = None
    loss = run()
    grads = torch.autograd.grad(loss, tr.opt.params, allow_unused=True)
    for (n, p), d, g in zip(tr.model.named_parameters(), direct, grads):
        assert g is not None, n
        tol = 1e-4 * float(g.abs().max()) + 1e-7
        assert float((d - g).abs().max()) <= tol, (n, float((d - g).abs().max()), tol)


def test_crops_kernel_equals_reference_corner_rule(cuda_lib):
    """scda_crop_regions against the reference's get_corner_from_center + slicing
    (tools/faster_rcnn_train_val.py:411-438, 528-557), centres at and beyond every border."""
    import torch
    from scda_b200.engine import _crops, crops_device, get_corner_from_center
    Hh, Ww, R = 512, 1024, 256
    g = torch.Generator().manual_seed(4)
    image = torch.randn(1, 3, Hh, Ww, generator=g)
    centers = np.array([[0.0, 0.0], [1023.9, 511.9], [127.99, 128.0], [128.0, 127.99], [896.0, 384.0],
                        [895.99, 383.99], [512.5, 256.5], [900.2, 10.7], [3.3, 500.1]], np.float32)
    want = _crops(image, get_corner_from_center(centers, R, Ww, Hh), R)
    got = crops_device(image.cuda(), torch.from_numpy(centers).cuda(), R, Ww, Hh)
    assert got.shape == (len(centers), 3, R, R)
    assert torch.equal(got.cpu(), want)


def test_stale_direct_gradient_is_zeroed_before_autograd_accumulates(cuda_lib):
    """FlatGradBucket.zero() leaves the big direct-written gradients un-zeroed (the tensor-core
    sinks overwrite them).  A gradient that reaches such a parameter through ordinary autograd — here the
    detector run through the plain fp32 torch graph under FlatAdam(tensor_core=True) — must not be added
    onto last step's values."""
    import torch
    from scda_b200.engine import FlatAdam
    from scda_b200.models.faster_rcnn.vgg_adver_expansion_cluster import vgg16
    cfg = _inputs.load_cfg()
    torch.manual_seed(0)
    model = vgg16(cfg=cfg["shared"]).cuda()
    opt = FlatAdam(model, 1e-4, tensor_core=True)
    model.eval()                                          # dropout off: the two steps compute the same gradient
    model._fp32_graph = True
    fc6 = model.classifier[0].weight                      # 102.8 M elements: never zeroed by zero()
    feat = torch.randn(1, 512, 16, 32, device="cuda")
    rois = torch.from_numpy(_inputs.rois_uniform(16, 2, img_w=512, img_h=256, wh=(16, 128))).cuda()
    grads = []
    for it in range(2):
        opt.zero_grad()
        fea, cls, loc = model.rcnn(feat, rois)
        (cls.sum() + loc.sum()).backward()
        opt.bucket.rebind()
        opt.bucket.settle()
        grads.append(fc6.grad.detach().clone())
    model._fp32_graph = False
    assert float(grads[0].abs().max()) > 0
    assert torch.allclose(grads[0], grads[1], rtol=1e-5, atol=1e-7), "second step accumulated onto stale gradients"


def test_prefetched_inputs_reach_the_graph_buffers(cuda_lib):
    """SCDATrainer.prefetch: the next iteration's pinned host inputs go up on the copy stream; iteration()
    called with the same tensors takes the device copies (no second H2D), with other tensors it copies
    itself.  Either way the graph's fixed input buffers hold exactly the inputs of THAT iteration."""
    import torch
    tr, cfg = _trainer(True)
    batches = []
    for seed in range(3):
        img, tgt, gts, info = _batch(seed)
        batches.append((img.cpu().pin_memory(), tgt.cpu().pin_memory(), gts.cpu().pin_memory(), info))
    tr.prefetch(batches[0][0], batches[0][2], batches[0][1])
    for it in range(5):                      # eager, capture + replay, replays
        img, tgt, gts, info = batches[it % 3]
        out = tr.iteration(cfg, img, info, gts, tgt)
        assert tr._prefetched is None
        if it in (0, 1, 3):                  # prefetch the next batch on some steps only
            nxt = batches[(it + 1) % 3]
            tr.prefetch(nxt[0], nxt[2], nxt[1])
        torch.cuda.synchronize()
        b = tr._static
        assert torch.equal(b['image'].cpu(), img) and torch.equal(b['target'].cpu(), tgt)
        assert torch.equal(b['gts'].cpu(), gts)
        assert bool(torch.isfinite(out['loss']))
