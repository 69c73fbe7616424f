"""Fused channels-last InstanceNorm(+activation) (csrc/norm_ops.cu) against
torch.nn.functional.instance_norm + activation (fp32), forward and backward, and the
reconstruction networks with the fused blocks against the same modules run through plain
torch.nn.  Tolerance: fp32 arithmetic in a different summation order -> 1e-4 relative
(north_star's bound for fp32 loss-path operators)."""
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-20))


@pytest.mark.parametrize("N,C,H,W", [(4, 128, 64, 64), (4, 64, 128, 128), (4, 32, 256, 256), (2, 8, 5, 7),
                                     (1, 256, 16, 16)])
@pytest.mark.parametrize("act", [None, "relu", "leaky_relu"])
def test_instance_norm_act(cuda_lib, N, C, H, W, act):
    import torch
    import torch.nn.functional as F
    from scda_b200.gan_ops import instance_norm_act
    g = torch.Generator(device="cuda").manual_seed(C + H)
    x = (torch.randn(N, C, H, W, device="cuda", generator=g) * 3 + 5).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(N, C, H, W, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    x1 = x.clone().requires_grad_(True)
    y = instance_norm_act(x1, act, 0.01)
    assert y.is_contiguous(memory_format=torch.channels_last)
    y.backward(dy)
    x2 = x.detach().double().contiguous().requires_grad_(True)      # float64 NCHW reference
    r = F.instance_norm(x2, eps=1e-5)
    r = F.relu(r) if act == "relu" else (F.leaky_relu(r, 0.01) if act == "leaky_relu" else r)
    r.backward(dy.double())
    assert _rel(y.double(), r.detach()) < 1e-4
    assert _rel(x1.grad.double(), x2.grad) < 1e-4
    # deterministic: same input, same bits
    y2 = instance_norm_act(x.clone(), act, 0.01)
    assert torch.equal(y2, y.detach())


@pytest.mark.parametrize("N,C,H,W", [(4, 128, 64, 64), (4, 64, 128, 128), (2, 8, 5, 7), (1, 4, 2, 3)])
def test_upsample_bilinear2x(cuda_lib, N, C, H, W):
    import torch
    import torch.nn.functional as F
    from scda_b200.gan_ops import upsample_bilinear2x
    g = torch.Generator(device="cuda").manual_seed(C + W)
    x = torch.randn(N, C, H, W, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(N, C, 2 * H, 2 * W, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    x1 = x.clone().requires_grad_(True)
    y = upsample_bilinear2x(x1)
    y.backward(dy)
    x2 = x.detach().contiguous().requires_grad_(True)
    r = F.interpolate(x2, scale_factor=2, mode='bilinear', align_corners=True)
    r.backward(dy.contiguous())
    assert y.shape == r.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert float((y - r).abs().max()) <= 1e-5 * float(r.abs().max())
    assert float((x1.grad - x2.grad).abs().max()) <= 1e-5 * float(x2.grad.abs().max())


def test_decoder_fused_equals_torch_modules(cuda_lib, monkeypatch):
    import copy
    import torch
    from scda_b200.engine import builder_gan
    from scda_b200.models.faster_rcnn import common_net
    torch.manual_seed(0)
    dis, dec, patch = builder_gan(4, 128, 256)
    for net in (dis, dec, patch):
        net.cuda().train()
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    ref = copy.deepcopy(dec)
    g = torch.Generator(device="cuda").manual_seed(1)
    xa = torch.randn(4, 128, 4096, device="cuda", generator=g)
    xb = torch.randn(4, 128, 4096, device="cuda", generator=g)
    torch.backends.cudnn.allow_tf32 = False
    # fp32 path (cuDNN convolutions + the fused fp32 InstanceNorm kernels); the tensor-core
    # path has its own test below
    monkeypatch.setattr(common_net, "conv_in_act_tc_supported", lambda x, conv: False)
    ya, yb = dec(xa, xb)
    (ya.square().mean() + yb.mean()).backward()
    monkeypatch.setattr(common_net, "_fused_ok", lambda x: False)
    monkeypatch.setattr(common_net, "upsample_supported", lambda x, s, m: False)
    monkeypatch.setattr(common_net, "conv1x1_tanh_supported", lambda x, c: False)
    ra, rb = ref(xa, xb)
    (ra.square().mean() + rb.mean()).backward()
    assert ya.shape == (4, 3, 256, 256)
    assert _rel(ya, ra) < 1e-3 and _rel(yb, rb) < 1e-3
    for (n, p), q in zip(dec.named_parameters(), ref.parameters()):
        # (the bias of a convolution that feeds an InstanceNorm has an exactly-zero true
        #  gradient: both sides hold rounding noise there, hence the absolute floor)
        assert float((p.grad - q.grad).abs().max()) <= 2e-3 * float(q.grad.abs().max()) + 1e-6, n
    # the discriminators accept the channels-last decoder output
    sa, sb = dis(ya.detach(), yb.detach())
    assert sa.shape == (4, 1024) and patch(xa).shape == (4, 512)


@pytest.mark.parametrize("cin,cout,stride,hw", [(128, 128, 1, 64), (3, 32, 2, 256), (64, 32, 1, 128), (32, 64, 2, 40)])
def test_conv2d_cl_matches_nn_conv2d(cuda_lib, cin, cout, stride, hw):
    """Conv2dCL = nn.Conv2d; only the bias gradient takes another route (scda_colsum_f32)."""
    import torch
    from scda_b200.models.faster_rcnn.common_net import Conv2dCL
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(cin + cout)
    ours = Conv2dCL(cin, cout, 3, stride, 1).cuda()
    ref = torch.nn.Conv2d(cin, cout, 3, stride, 1).cuda()
    ref.load_state_dict(ours.state_dict())
    x = torch.randn(4, cin, hw, hw, device="cuda").contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = ours(xa), ref(xb)
    assert torch.allclose(ya, yb, rtol=1e-4, atol=1e-5)
    g = torch.randn_like(ya)
    ya.backward(g)
    yb.backward(g)
    assert torch.allclose(xa.grad, xb.grad, rtol=1e-3, atol=1e-4)
    assert torch.allclose(ours.weight.grad, ref.weight.grad, rtol=1e-3, atol=1e-2)
    # bias gradient = sum of g over (N, H, W): fp32 sums of ~1e5 terms in a different order
    assert torch.allclose(ours.bias.grad, ref.bias.grad, rtol=1e-3, atol=2e-2)
    exact = g.double().sum((0, 2, 3))
    assert float((ours.bias.grad.double() - exact).abs().max()) <= float((ref.bias.grad.double() - exact).abs().max()) + 1e-2


def test_decoder_tensor_core_path(cuda_lib, monkeypatch):
    """The decoder with its 64-multiple 3x3 convolutions on the tcgen05 halo kernel (bf16
    operands, fp32 accumulate, conv + InstanceNorm + activation as one autograd node) against the
    same modules in plain fp32 torch.  Tolerance: bf16 operand rounding (2^-9 relative per
    operand) through 8 convolutions, each re-normalised by its InstanceNorm -> a few 1e-2 of the
    output range; gradients by direction and size."""
    import copy
    import torch
    from scda_b200 import gan_ops
    from scda_b200.engine import builder_gan
    from scda_b200.models.faster_rcnn import common_net
    assert gan_ops.TC_GAN
    torch.manual_seed(0)
    _, dec, _ = builder_gan(4, 128, 256)
    dec.cuda().train()
    for m in dec.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    ref = copy.deepcopy(dec)
    g = torch.Generator(device="cuda").manual_seed(1)
    xa = torch.randn(4, 128, 4096, device="cuda", generator=g).requires_grad_(True)
    xb = torch.randn(4, 128, 4096, device="cuda", generator=g)
    xr = xa.detach().clone().requires_grad_(True)
    torch.backends.cudnn.allow_tf32 = False
    calls = []
    real = common_net.conv_in_act_tc

    def spy(*a, **k):
        calls.append(1)
        return real(*a, **k)
    monkeypatch.setattr(common_net, "conv_in_act_tc", spy)
    ya, yb = dec(xa, xb)
    assert len(calls) == 16, "6 residual convolutions + 2 up-sampling convolutions per decoder"
    (ya.square().mean() + yb.mean()).backward()
    monkeypatch.setattr(common_net, "conv_in_act_tc_supported", lambda x, conv: False)
    monkeypatch.setattr(common_net, "_fused_ok", lambda x: False)
    monkeypatch.setattr(common_net, "upsample_supported", lambda x, s, m: False)
    monkeypatch.setattr(common_net, "conv1x1_tanh_supported", lambda x, c: False)
    ra, rb = ref(xr, xb)
    (ra.square().mean() + rb.mean()).backward()
    assert _rel(ya, ra) < 5e-2 and _rel(yb, rb) < 5e-2
    assert float((ya - ra).pow(2).mean().sqrt() / ra.pow(2).mean().sqrt()) < 2e-2

    def cos(a, b):
        return float((a * b).sum() / (a.norm() * b.norm()).clamp(min=1e-30))
    assert cos(xa.grad, xr.grad) > 0.99 and 0.9 < float(xa.grad.norm() / xr.grad.norm()) < 1.1
    for (n, p), q in zip(dec.named_parameters(), ref.parameters()):
        if n.endswith("weight"):
            assert cos(p.grad, q.grad) > 0.98, (n, cos(p.grad, q.grad))
            assert 0.9 < float(p.grad.norm() / q.grad.norm()) < 1.1, n


@pytest.mark.parametrize("dy_bf16,dx_bf16", [(False, True), (True, True), (True, False)])
def test_instance_norm_bf16_io(cuda_lib, dy_bf16, dx_bf16):
    """scda_instnorm_act_{fwd,bwd}_nhwc with bf16 outputs / incoming gradients = the fp32 kernels
    followed / preceded by a round to bf16."""
    import torch
    from scda_b200._lib import check, load, stream_ptr
    N, H, W, C = 4, 32, 32, 128
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(N, H, W, C, device="cuda", generator=g) * 2 + 1
    dy = torch.randn(N, H, W, C, device="cuda", generator=g)
    lib = load()
    wsb = lib.scda_instnorm_workspace_bytes(N, H * W, C) + 8 * N * C
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    mean, rstd = torch.empty(N, C, device="cuda"), torch.empty(N, C, device="cuda")
    y32, y16 = torch.empty_like(x), torch.empty(N, H, W, C, dtype=torch.bfloat16, device="cuda")
    for y, code in ((y32, 0), (y16, 1)):
        check(lib.scda_instnorm_act_fwd_nhwc(N, H * W, C, x.data_ptr(), y.data_ptr(), code, mean.data_ptr(),
                                             rstd.data_ptr(), 1e-5, 2, 0.01, ws.data_ptr(), wsb, stream_ptr()), "fwd")
    assert torch.equal(y16, y32.bfloat16())
    dyi = dy.bfloat16() if dy_bf16 else dy
    dx_ref = torch.empty_like(x)
    check(lib.scda_instnorm_act_bwd_nhwc(N, H * W, C, x.data_ptr(), dyi.float().data_ptr(), 0, mean.data_ptr(),
                                         rstd.data_ptr(), dx_ref.data_ptr(), 0, 2, 0.01, ws.data_ptr(), wsb,
                                         stream_ptr()), "bwd ref")
    dx = torch.empty(N, H, W, C, dtype=torch.bfloat16 if dx_bf16 else torch.float32, device="cuda")
    check(lib.scda_instnorm_act_bwd_nhwc(N, H * W, C, x.data_ptr(), dyi.data_ptr(), 1 if dy_bf16 else 0,
                                         mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(), 1 if dx_bf16 else 0, 2, 0.01,
                                         ws.data_ptr(), wsb, stream_ptr()), "bwd")
    assert torch.equal(dx, dx_ref.bfloat16() if dx_bf16 else dx_ref)


@pytest.mark.parametrize("cout,hw,bias", [(3, 256, True), (3, 40, False), (1, 64, True), (4, 33, True)])
def test_decoder_head_conv1x1_tanh(cuda_lib, cout, hw, bias):
    """scda_conv1x1_tanh_{fwd,bwd} = nn.ConvTranspose2d(32, cout, 1) + nn.Tanh (fp32, 1e-5 relative)."""
    import torch
    from scda_b200.gan_ops import conv1x1_tanh, conv1x1_tanh_supported
    torch.manual_seed(cout * 100 + hw)
    convt = torch.nn.ConvTranspose2d(32, cout, 1, bias=bias).cuda()
    x = torch.randn(4, 32, hw, hw, device="cuda").contiguous(memory_format=torch.channels_last)
    dy = torch.randn(4, cout, hw, hw, device="cuda")
    assert conv1x1_tanh_supported(x, convt)
    xa = x.clone().requires_grad_(True)
    y = conv1x1_tanh(xa, convt)
    assert getattr(y, "_scda_tanh_applied", False) and y.shape == (4, cout, hw, hw)
    y.backward(dy)
    got = [xa.grad.clone(), convt.weight.grad.clone()] + ([convt.bias.grad.clone()] if bias else [])
    convt.zero_grad()
    xr = x.clone().double().requires_grad_(True)
    ref_mod = torch.nn.ConvTranspose2d(32, cout, 1, bias=bias).cuda().double()
    ref_mod.load_state_dict({k: v.double() for k, v in convt.state_dict().items()})
    r = torch.tanh(ref_mod(xr))
    r.backward(dy.double())
    want = [xr.grad, ref_mod.weight.grad] + ([ref_mod.bias.grad] if bias else [])
    assert _rel(y.double(), r.detach()) < 1e-5
    for a, b in zip(got, want):
        assert _rel(a.double(), b) < 2e-5
    # deterministic
    y2 = conv1x1_tanh(x.clone().requires_grad_(True), convt)
    assert torch.equal(y2, y.detach())
