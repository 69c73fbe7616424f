"""The discriminators on hand-written kernels (scda_b200/disc_ops.py, csrc/disc_ops.cu, the stride-2 form of
csrc/conv_halo.cu / tc_wgrad_kernel) against the torch modules they replace (cuDNN fp32, TF32 off):
GAN_dis_AE / GAN_dis_AE_patch of the reference, faster_rcnn_adver_expansion_reweight_cluster.py:270-333,
common_net.py:205-261."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp(min=1e-30))


def _no_tf32():
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("N,H,W,C,Cout", [(4, 64, 64, 32, 64), (2, 32, 32, 64, 128), (4, 64, 64, 128, 256),
                                          (4, 32, 32, 256, 512), (4, 16, 16, 512, 512), (1, 48, 80, 64, 96)])
def test_conv_s2_fwd_dgrad_wgrad(cuda_lib, N, H, W, C, Cout):
    """stride-2 3x3 convolution, its data gradient (with the LeakyReLU mask) and its weight gradient against
    torch fp32 on the same bf16-rounded operands"""
    import torch
    import torch.nn.functional as F
    from scda_b200 import disc_ops
    _no_tf32()
    g = torch.Generator(device="cuda").manual_seed(N * H + C)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).bfloat16()
    w = torch.nn.Parameter((torch.randn(Cout, C, 3, 3, device="cuda", generator=g) / (9 * C) ** 0.5)
                           .contiguous(memory_format=torch.channels_last))
    b = torch.randn(Cout, device="cuda", generator=g)
    wd = disc_ops.s2_weights(w)
    wr = w.detach().bfloat16().float()
    xn = x.float().permute(0, 3, 1, 2)
    ref = F.leaky_relu(F.conv2d(xn, wr, b, stride=2, padding=1), 0.01).permute(0, 2, 3, 1)
    y = disc_ops.conv_s2(x, wd, b, disc_ops.LEAKY, 0.01)
    assert y.shape == ref.shape
    assert _rel(y.float(), ref) < 6e-3, _rel(y.float(), ref)
    y32 = disc_ops.conv_s2(x, wd, None, 0, out_dtype=torch.float32)
    ref32 = F.conv2d(xn, wr, None, stride=2, padding=1).permute(0, 2, 3, 1)
    assert _rel(y32, ref32) < 1e-4, _rel(y32, ref32)
    dy = torch.randn(ref.shape, device="cuda", generator=g).bfloat16()
    dref = torch.nn.grad.conv2d_input(xn.shape, wr, dy.float().permute(0, 3, 1, 2), stride=2, padding=1)
    dref = dref.permute(0, 2, 3, 1)
    dx = disc_ops.conv_s2_dgrad(dy, wd, C)
    assert dx.shape == x.shape and _rel(dx.float(), dref) < 6e-3, _rel(dx.float(), dref)
    dxm = disc_ops.conv_s2_dgrad(dy, wd, C, mask_src=x, slope=0.01)
    drefm = dref * torch.where(x.float() > 0, torch.ones_like(dref), torch.full_like(dref, 0.01))
    assert _rel(dxm.float(), drefm) < 6e-3
    dw = disc_ops.conv_s2_wgrad(x, dy)
    wref = torch.nn.grad.conv2d_weight(xn, w.shape, dy.float().permute(0, 3, 1, 2), stride=2, padding=1)
    assert _rel(dw.permute(0, 3, 1, 2), wref) < 1e-4, _rel(dw.permute(0, 3, 1, 2), wref)


def _dis_nets(seed=0):
    import torch
    from scda_b200.engine import builder_gan
    torch.manual_seed(seed)
    dis, dec, patch = builder_gan(4, 128, 256)
    dis, patch = dis.cuda(), patch.cuda()
    with torch.no_grad():
        for net in (dis, patch):
            for n, p in net.named_parameters():
                if p.dim() == 4:
                    p.mul_(5.0)                    # N(0, 0.02) init leaves the activations tiny: exercise the kernels
                elif n.endswith("bias"):
                    p.normal_(0, 0.1)
                elif n.endswith("weight"):
                    p.uniform_(0.5, 1.5)           # BatchNorm gamma
    return dis.train(), patch.train()


def _grads(net):
    return {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}


def _ste(x):
    """round to bf16 in the forward pass, identity in the backward pass"""
    return x + (x.bfloat16().float() - x).detach()


def _emulated_image_dis(seq, x):
    """plain torch fp32 graph on the SAME rounded operands the kernels see: layer 1 and the head in fp32 weights,
    layers 2 / 3 with bf16-rounded weights, every activation rounded to bf16 after its LeakyReLU"""
    import torch.nn.functional as F
    h = x
    for i in range(3):
        c = seq[i].model[0]
        w = c.weight if i == 0 else _ste(c.weight)
        h = _ste(F.leaky_relu(F.conv2d(h, w, c.bias, stride=2, padding=1), 0.01))
    o = F.conv2d(h, seq[3].weight, seq[3].bias)
    return o.reshape(o.size(0), -1)


@pytest.mark.parametrize("strided", [False, True])
def test_image_discriminator_matches_torch_modules(cuda_lib, monkeypatch, strided):
    import torch
    from scda_b200 import disc_ops
    _no_tf32()
    dis, _ = _dis_nets()
    g = torch.Generator(device="cuda").manual_seed(1)
    xa = torch.randn(4, 3, 256, 256, device="cuda", generator=g)
    xb = torch.randn(4, 3, 256, 256, device="cuda", generator=g)
    if strided:             # channels-last reconstructions, as the decoder head hands them over
        xa = xa.contiguous(memory_format=torch.channels_last)
        xb = xb.contiguous(memory_format=torch.channels_last)
    xa.requires_grad_(True)
    xb.requires_grad_(True)
    wgt = torch.randn(4, 1024, device="cuda", generator=g)

    def run(mode):
        dis.zero_grad()
        xa.grad = xb.grad = None
        if mode == "emulated":
            oa, ob = _emulated_image_dis(dis.model_A, xa), _emulated_image_dis(dis.model_B, xb)
        else:
            oa, ob = dis(xa, xb)
        ((oa * wgt).sum() + (ob * wgt.flip(1)).sum()).backward()
        return oa.detach(), ob.detach(), _grads(dis), xa.grad.clone(), xb.grad.clone()

    assert disc_ops.image_dis_supported(dis.model_A, xa)
    oa, ob, gk, gxa, gxb = run("kernels")
    ea, eb, ge, exa, exb = run("emulated")
    monkeypatch.setattr(disc_ops, "image_dis_supported", lambda seq, x: False)
    ra, rb, gr, rxa, rxb = run("modules")
    assert oa.shape == ra.shape == (4, 1024)
    # against the torch modules (cuDNN fp32): bf16 activations -> ~1e-2; the gradients carry the LeakyReLU masks
    # of bf16-rounded activations (a mask flip moves dx by 0.99 g: the sqrt(noise) law of
    # test_tc_detector_gpu.py::test_detector_stages_match_fp32_graph_x3)
    assert _rel(oa, ra) < 1.5e-2 and _rel(ob, rb) < 1.5e-2, (_rel(oa, ra), _rel(ob, rb))
    assert set(gk) == set(gr) == set(ge)
    bad = {n: _rel(gk[n], gr[n]) for n in gr if _rel(gk[n], gr[n]) > 0.15}
    assert not bad, bad
    assert _rel(gxa, rxa) < 0.15 and _rel(gxb, rxb) < 0.15, (_rel(gxa, rxa), _rel(gxb, rxb))
    # against the same-operand graph: only accumulation order and the bf16 rounding of the back-propagated
    # gradients differ
    assert _rel(oa, ea) < 3e-3 and _rel(ob, eb) < 3e-3, (_rel(oa, ea), _rel(ob, eb))
    bad = {n: _rel(gk[n], ge[n]) for n in ge if _rel(gk[n], ge[n]) > 2e-2}
    assert not bad, bad
    assert _rel(gxa, exa) < 2e-2 and _rel(gxb, exb) < 2e-2, (_rel(gxa, exa), _rel(gxb, exb))


def test_first_layer_and_head_are_fp32_exact(cuda_lib):
    """the direct first-layer and head kernels compute in fp32: tight against torch"""
    import torch
    import torch.nn.functional as F
    from scda_b200._lib import check, load, stream_ptr
    _no_tf32()
    lib = load()
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(2, 3, 64, 96, device="cuda", generator=g)
    w = torch.randn(32, 3, 3, 3, device="cuda", generator=g) * 0.2
    b = torch.randn(32, device="cuda", generator=g) * 0.1
    wk = w.permute(0, 2, 3, 1).contiguous()
    y = torch.empty(2, 32, 48, 32, device="cuda")
    st = stream_ptr(x.device)
    check(lib.scda_disc_l1_fwd(2, 64, 96, x.data_ptr(), *x.stride(), wk.data_ptr(), b.data_ptr(), 0.01, y.data_ptr(), 1,
                               st), "l1")
    ref = F.leaky_relu(F.conv2d(x, w, b, stride=2, padding=1), 0.01).permute(0, 2, 3, 1)
    assert _rel(y, ref) < 1e-6
    gy = torch.randn_like(y)
    dw, db = torch.empty(32, 3, 3, 3, device="cuda"), torch.empty(32, device="cuda")
    dx = torch.empty(2, 64, 96, 3, device="cuda")
    wsb = lib.scda_disc_l1_workspace_bytes(2, 64, 96)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    check(lib.scda_disc_l1_bwd(2, 64, 96, x.data_ptr(), *x.stride(), wk.data_ptr(), gy.data_ptr(), 1, dw.data_ptr(),
                               db.data_ptr(), dx.data_ptr(), 0, ws.data_ptr(), wsb, st), "l1b")
    gn = gy.permute(0, 3, 1, 2)
    assert _rel(dw.permute(0, 3, 1, 2), torch.nn.grad.conv2d_weight(x, w.shape, gn, stride=2, padding=1)) < 1e-5
    assert _rel(db, gn.sum((0, 2, 3))) < 1e-5
    assert _rel(dx.permute(0, 3, 1, 2), torch.nn.grad.conv2d_input(x.shape, w, gn, stride=2, padding=1)) < 1e-5


def test_patch_discriminator_matches_torch_modules(cuda_lib, monkeypatch):
    import torch
    from scda_b200 import disc_ops
    _no_tf32()
    _, patch = _dis_nets(1)
    g = torch.Generator(device="cuda").manual_seed(3)
    feats = torch.randn(4, 128, 4096, device="cuda", generator=g).clamp(min=0)
    wgt = torch.randn(4, 512, device="cuda", generator=g)

    def run():
        patch.zero_grad()
        for m in patch.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.reset_running_stats()
        out = patch(feats)
        (out * wgt).sum().backward()
        stats = {n: b.detach().clone() for n, b in patch.named_buffers()}
        return out.detach(), _grads(patch), stats

    ok, gk, sk = run()
    monkeypatch.setattr(disc_ops, "patch_dis_supported", lambda seq, x: False)
    rk, gr, sr = run()
    assert ok.shape == rk.shape == (4, 512)
    assert _rel(ok, rk) < 1e-2, _rel(ok, rk)
    assert set(gk) == set(gr)
    # gradients: bf16 operands under LeakyReLU masks (see the image discriminator test); the convolution and
    # BatchNorm kernels themselves are pinned to 1e-4 above / below
    bad = {n: _rel(gk[n], gr[n]) for n in gr if _rel(gk[n], gr[n]) > 0.15}
    assert not bad, bad
    for n in sr:
        if sr[n].dtype.is_floating_point:
            assert _rel(sk[n], sr[n]) < 1e-2, (n, _rel(sk[n], sr[n]))
        else:
            assert torch.equal(sk[n], sr[n]), n


def test_bn_lrelu_kernels_match_torch(cuda_lib):
    import torch
    import torch.nn.functional as F
    from scda_b200._lib import check, load, stream_ptr
    lib = load()
    g = torch.Generator(device="cuda").manual_seed(4)
    P, C = 4096, 256
    x = (torch.randn(P, C, device="cuda", generator=g) * 2 + 0.5).requires_grad_(True)
    gamma = torch.rand(C, device="cuda", generator=g).add_(0.5).requires_grad_(True)
    beta = torch.randn(C, device="cuda", generator=g).requires_grad_(True)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    rm2, rv2 = rm.clone(), rv.clone()
    ref = F.leaky_relu(F.batch_norm(x, rm2, rv2, gamma, beta, True, 0.1, 1e-5), 0.01)
    dy = torch.randn(P, C, device="cuda", generator=g)
    ref.backward(dy)
    mean, rstd = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    y = torch.empty(P, C, device="cuda")
    st = stream_ptr(x.device)
    wsb = lib.scda_bn_workspace_bytes(P, C)
    ws = torch.full((wsb,), 0xAB, dtype=torch.uint8, device="cuda")     # contents irrelevant
    check(lib.scda_bn_lrelu_fwd(P, C, x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-5, 0.01, 0.1, rm.data_ptr(),
                                rv.data_ptr(), mean.data_ptr(), rstd.data_ptr(), y.data_ptr(), 1, ws.data_ptr(), wsb,
                                st), "bn")
    assert _rel(y, ref) < 1e-5 and _rel(rm, rm2) < 1e-5 and _rel(rv, rv2) < 1e-5
    dx, dg, db = torch.empty(P, C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    check(lib.scda_bn_lrelu_bwd(P, C, x.data_ptr(), dy.data_ptr(), 1, gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(),
                                rstd.data_ptr(), 0.01, dx.data_ptr(), 1, dg.data_ptr(), db.data_ptr(), 0, ws.data_ptr(),
                                wsb, st), "bnb")
    assert _rel(dx, x.grad) < 1e-4 and _rel(dg, gamma.grad) < 1e-4 and _rel(db, beta.grad) < 1e-4
