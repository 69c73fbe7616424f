"""The halo form of the 3x3 convolution (csrc/conv_halo.cu: input tile staged once, nine taps read
through shifted UMMA descriptors) under every tile plan, forward and data gradient, against a
plain PyTorch fp32 convolution of the same bf16-rounded operands and against the per-tap kernel.
Tolerance as in test_tc_gpu.py: fp32 accumulation, bf16 output rounding."""
import pytest

pytestmark = pytest.mark.gpu

PLANS = [(1, 64, 1), (1, 64, 2), (1, 128, 1), (1, 128, 2), (1, 0, 0), (0, 0, 0)]
SHAPES = [(1, 16, 16, 64, 64), (1, 32, 64, 128, 128), (2, 32, 64, 64, 256), (1, 24, 40, 64, 192),
          (1, 20, 24, 192, 128), (3, 16, 8, 128, 64), (1, 64, 128, 256, 512), (1, 48, 72, 64, 64),
          (2, 32, 32, 64, 32), (1, 64, 64, 128, 32)]      # 32 output channels: the decoder's last 3x3 layer


def _close(out, ref, rtol=1e-2):
    out, ref = out.float(), ref.float()
    tol = rtol * ref.abs() + rtol * ref.pow(2).mean().sqrt()
    bad = (out - ref).abs() > tol
    assert not bool(bad.any()), "mismatch: %d / %d, max abs err %g (ref rms %g)" % (
        int(bad.sum()), bad.numel(), float((out - ref).abs().max()), float(ref.pow(2).mean().sqrt()))


@pytest.fixture
def plan_reset(cuda_lib):
    from scda_b200 import tc
    yield tc
    tc.set_conv_plan(1, 0, 0)


@pytest.mark.parametrize("NB,H,W,Cin,Cout", SHAPES)
def test_forward_all_plans(plan_reset, NB, H, W, Cin, Cout):
    import torch
    import torch.nn.functional as F
    tc = plan_reset
    g = torch.Generator(device="cuda").manual_seed(H * W + Cin + Cout)
    x = torch.randn(NB, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, 3, 3, Cin, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=1)
    ref = ref.permute(0, 2, 3, 1)
    for halo, bn, sub in PLANS:
        tc.set_conv_plan(halo, bn, sub)
        _close(tc.conv3x3_nhwc(x, w, bias), ref)
        _close(tc.conv3x3_nhwc(x, w, bias, relu=True, out_dtype=torch.float32), ref.clamp(min=0), rtol=2e-3)


@pytest.mark.parametrize("NB,H,W,Cin,Cout", SHAPES)
def test_dgrad_all_plans(plan_reset, NB, H, W, Cin, Cout):
    import torch
    tc = plan_reset
    g = torch.Generator(device="cuda").manual_seed(H + W + Cin * Cout)
    x = torch.randn(NB, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, 3, 3, Cin, device="cuda", generator=g) / (9 * Cout) ** 0.5).bfloat16()
    dy = torch.randn(NB, H, W, Cout, device="cuda", generator=g).bfloat16()
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.grad.conv2d_input((NB, Cin, H, W), w.float().permute(0, 3, 1, 2),
                                     dy.float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    refm = torch.where(x.float() > 0, ref, torch.zeros_like(ref))
    for halo, bn, sub in PLANS:
        if not halo and Cout % 64:
            continue                       # the per-tap form reduces over whole 64-channel blocks
        tc.set_conv_plan(halo, bn, sub)
        _close(tc.conv3x3_dgrad_nhwc(dy, w), ref)
        _close(tc.conv3x3_dgrad_nhwc(dy, w, mask_src=x), refm)


def test_halo_equals_per_tap_on_integers(plan_reset):
    """Small-integer operands make every product and partial sum exact in fp32, so the two kernels
    (different summation orders) must agree BIT FOR BIT."""
    import torch
    tc = plan_reset
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randint(-3, 4, (1, 32, 48, 128), device="cuda", generator=g).bfloat16()
    w = torch.randint(-2, 3, (128, 3, 3, 128), device="cuda", generator=g).bfloat16()
    tc.set_conv_plan(0, 0, 0)
    y0 = tc.conv3x3_nhwc(x, w, out_dtype=torch.float32)
    d0 = tc.conv3x3_dgrad_nhwc(x, w, out_dtype=torch.float32)
    for bn, sub in ((64, 1), (64, 2), (128, 1), (128, 2)):
        tc.set_conv_plan(1, bn, sub)
        assert torch.equal(tc.conv3x3_nhwc(x, w, out_dtype=torch.float32), y0)
        assert torch.equal(tc.conv3x3_dgrad_nhwc(x, w, out_dtype=torch.float32), d0)


def test_backbone_full_size_persistent_tiles(plan_reset):
    """conv2_2 at the benchmark resolution (128 -> 128 @ 256 x 512): many tiles per CTA, both
    accumulator buffers and every pipeline slot wrap; exact-integer comparison with the per-tap
    kernel."""
    import torch
    tc = plan_reset
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randint(-2, 3, (1, 256, 512, 128), device="cuda", generator=g).bfloat16()
    w = torch.randint(-1, 2, (128, 3, 3, 128), device="cuda", generator=g).bfloat16()
    tc.set_conv_plan(0, 0, 0)
    y0 = tc.conv3x3_nhwc(x, w, out_dtype=torch.float32)
    for bn, sub in ((128, 2), (64, 1)):
        tc.set_conv_plan(1, bn, sub)
        assert torch.equal(tc.conv3x3_nhwc(x, w, out_dtype=torch.float32), y0)


@pytest.fixture
def pair_reset(plan_reset):
    yield plan_reset
    plan_reset.set_conv_pair(-1)


PAIR_SHAPES = [(1, 16, 16, 64, 64), (1, 32, 64, 128, 128), (2, 32, 64, 64, 256), (1, 24, 40, 64, 192),
               (3, 16, 8, 128, 64), (1, 64, 128, 256, 512), (1, 16, 24, 64, 256), (1, 48, 72, 64, 64)]
# (1, 16, 24, ...): 3 pixel tiles per row, 6 in all; (3, 16, 8, ...): 3 tiles = an odd count, phantom second tile


@pytest.mark.timeout(120)
@pytest.mark.parametrize("NB,H,W,Cin,Cout", PAIR_SHAPES)
def test_pair_form_forward_and_dgrad(pair_reset, NB, H, W, Cin, Cout):
    """CTA-pair form (cta_group::2, half a weight tile per SM) under every legal N tile against the fp32
    convolution of the same operands, and bit for bit against the lone-CTA plan on small integers."""
    import torch
    import torch.nn.functional as F
    tc = pair_reset
    g = torch.Generator(device="cuda").manual_seed(H * W + Cin + Cout)
    x = torch.randn(NB, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, 3, 3, Cin, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g)
    dy = torch.randn(NB, H, W, Cout, device="cuda", generator=g).bfloat16()
    xi = torch.randint(-3, 4, (NB, H, W, Cin), device="cuda", generator=g).bfloat16()
    wi = torch.randint(-2, 3, (Cout, 3, 3, Cin), device="cuda", generator=g).bfloat16()
    dyi = torch.randint(-3, 4, (NB, H, W, Cout), device="cuda", generator=g).bfloat16()
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=1).permute(0, 2, 3, 1)
    dref = torch.nn.grad.conv2d_input((NB, Cin, H, W), w.float().permute(0, 3, 1, 2),
                                      dy.float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    tc.set_conv_pair(0)
    tc.set_conv_plan(1, 0, 0)
    y0 = tc.conv3x3_nhwc(xi, wi, out_dtype=torch.float32)
    d0 = tc.conv3x3_dgrad_nhwc(dyi, wi, out_dtype=torch.float32)
    tc.set_conv_pair(1)
    for bn in (0, 64, 128, 256):
        tc.set_conv_plan(1, bn, 0)
        _close(tc.conv3x3_nhwc(x, w, bias), ref)
        _close(tc.conv3x3_nhwc(x, w, bias, relu=True, out_dtype=torch.float32), ref.clamp(min=0), rtol=2e-3)
        assert torch.equal(tc.conv3x3_nhwc(xi, wi, out_dtype=torch.float32), y0)
        _close(tc.conv3x3_dgrad_nhwc(dy, w), dref)
        _close(tc.conv3x3_dgrad_nhwc(dy, w, mask_src=x), torch.where(x.float() > 0, dref, torch.zeros_like(dref)))
        assert torch.equal(tc.conv3x3_dgrad_nhwc(dyi, wi, out_dtype=torch.float32), d0)


@pytest.mark.timeout(120)
def test_pair_form_full_size_persistent_tiles(pair_reset):
    """conv3_2 at the benchmark resolution (256 -> 256 @ 128 x 256): 128 tile groups over 74 SM pairs, so
    pipeline slots and both accumulator buffers wrap; exact-integer comparison with the lone-CTA plan."""
    import torch
    tc = pair_reset
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randint(-2, 3, (1, 128, 256, 256), device="cuda", generator=g).bfloat16()
    w = torch.randint(-1, 2, (256, 3, 3, 256), device="cuda", generator=g).bfloat16()
    tc.set_conv_pair(0)
    tc.set_conv_plan(1, 0, 0)
    y0 = tc.conv3x3_nhwc(x, w, out_dtype=torch.float32)
    d0 = tc.conv3x3_dgrad_nhwc(x, w, out_dtype=torch.float32)
    tc.set_conv_pair(1)
    for bn in (64, 128, 256):
        tc.set_conv_plan(1, bn, 0)
        assert torch.equal(tc.conv3x3_nhwc(x, w, out_dtype=torch.float32), y0)
        assert torch.equal(tc.conv3x3_dgrad_nhwc(x, w, out_dtype=torch.float32), d0)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("NB,Cin,H,W", [(1, 3, 32, 128), (2, 3, 20, 24), (1, 3, 17, 33), (1, 1, 16, 16), (1, 2, 9, 130),
                                         (1, 3, 128, 256)])
def test_first_convolution_from_the_nchw_image(cuda_lib, NB, Cin, H, W):
    """conv1_1 straight from the fp32 NCHW image (csrc/conv_first.cu: im2col in shared memory, two K = 16 MMAs per
    128 pixels) against torch's fp32 convolution of the same bf16-rounded operands; ragged pixel counts, image
    borders, 1-3 input channels.  Exact on small integers."""
    import torch
    import torch.nn.functional as F
    from scda_b200 import tc
    g = torch.Generator(device="cuda").manual_seed(NB * H * W + Cin)
    x = torch.randn(NB, Cin, H, W, device="cuda", generator=g)
    w = (torch.randn(64, 3, 3, Cin, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(64, device="cuda", generator=g)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(x.bfloat16().float(), w.float().permute(0, 3, 1, 2), bias, padding=1).permute(0, 2, 3, 1)
    _close(tc.conv3x3_first_nchw(x, w, bias, relu=False), ref)
    _close(tc.conv3x3_first_nchw(x, w, bias, relu=True), ref.clamp(min=0))
    xi = torch.randint(-4, 5, (NB, Cin, H, W), device="cuda", generator=g).float()
    wi = torch.randint(-3, 4, (64, 3, 3, Cin), device="cuda", generator=g).bfloat16()
    bi = torch.randint(-5, 6, (64,), device="cuda", generator=g).float()
    refi = F.conv2d(xi, wi.float().permute(0, 3, 1, 2), bi, padding=1).permute(0, 2, 3, 1)
    assert torch.equal(tc.conv3x3_first_nchw(xi, wi, bi, relu=False).float(), refi.bfloat16().float())
