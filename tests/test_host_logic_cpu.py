"""Host-side logic that needs no GPU: the split plan of the weight-gradient kernels, the
sequential fallback of the stream-pair helper, the CPU fallbacks of the reconstruction-network
modules (every fused path is CUDA-only and must hand CPU tensors to the plain torch layers), and
the 1x1 ConvTranspose evaluated as a convolution."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("three", [False, True])
def test_wgrad_split_plan_covers_every_slab(three, monkeypatch):
    """scda_conv3x3_wgrad_bf16_nhwc rejects a plan with an empty slab (every slab must be written):
    the Python planner must only produce plans whose last range is non-empty."""
    from scda_b200 import tc
    monkeypatch.setattr(tc, "WGRAD3", three)
    for NB, H, W, Cin, Cout in [(1, 512, 1024, 64, 64), (1, 256, 512, 128, 128), (1, 128, 256, 256, 256),
                                (1, 64, 128, 512, 512), (1, 32, 64, 512, 512), (4, 64, 64, 128, 128),
                                (4, 256, 256, 64, 32), (1, 8, 16, 64, 128), (2, 16, 8, 64, 64)]:
        for target in (9, 148, 296, 1000):
            tiles = NB * H * W // 128
            splits = tc._wgrad_splits(NB, H, W, Cin, Cout, target)
            assert 1 <= splits <= tiles
            per = -(-tiles // splits)
            assert -(-tiles // per) == splits, (NB, H, W, Cin, Cout, target, splits)


def test_run_pair_is_sequential_without_streams():
    from scda_b200 import gan_ops
    calls = []
    a, b = gan_ops.run_pair(lambda: calls.append("a") or 1, lambda: calls.append("b") or 2)
    assert (a, b) == (1, 2) and calls == ["a", "b"]


def test_reconstruction_modules_fall_back_on_cpu():
    """CPU tensors never reach a fused CUDA path: the decoder and both discriminators run as plain
    torch modules (this is the path oracle/model_cpu.py times) and give the reference's shapes."""
    from scda_b200 import gan_ops
    from scda_b200.engine import builder_gan
    from scda_b200.models.faster_rcnn import common_net
    torch.manual_seed(0)
    dis, dec, patch = builder_gan(cluster_num=2, threshold=128, recon_size=256)
    x = torch.randn(2, 128, 4096)
    conv = dec.decode_A[1].model[0]
    assert not gan_ops.conv_in_act_tc_supported(x.view(2, 128, 64, 64), conv)
    assert not gan_ops.conv_bias_supported(x.view(2, 128, 64, 64), conv)
    assert not gan_ops.conv1x1_tanh_supported(torch.randn(2, 32, 8, 8), dec.decode_A[-2])
    for net in (dis, dec, patch):
        net.eval()
    with torch.no_grad():
        ya, yb = dec(x, x)
    assert ya.shape == (2, 3, 256, 256) and float(ya.abs().max()) <= 1.0        # tanh applied exactly once
    head, tanh = dec.decode_A[-2], dec.decode_A[-1]
    assert isinstance(head, common_net.ConvTranspose1x1) and isinstance(tanh, common_net.TanhAfterHead)
    with torch.no_grad():
        sa, sb = dis(ya, yb)
    assert sa.shape == (2, 1024) and sb.shape == (2, 1024)


def test_tanh_after_head_passes_tagged_tensors_through():
    from scda_b200.models.faster_rcnn.common_net import TanhAfterHead
    t = TanhAfterHead()
    x = torch.tensor([0.5, -2.0])
    assert torch.equal(t(x), torch.tanh(x))
    y = x.clone()
    y._scda_tanh_applied = True
    assert t(y) is y


def test_conv_transpose_1x1_equals_nn_conv_transpose():
    """ConvTranspose1x1 keeps nn.ConvTranspose2d's parameters ([in, out, 1, 1]) and evaluates
    the equivalent convolution: same output, same gradients."""
    from scda_b200.models.faster_rcnn.common_net import ConvTranspose1x1
    torch.manual_seed(1)
    ours = ConvTranspose1x1(32, 3, kernel_size=1, stride=1, padding=0)
    ref = torch.nn.ConvTranspose2d(32, 3, kernel_size=1, stride=1, padding=0)
    ref.load_state_dict(ours.state_dict())
    x = torch.randn(2, 32, 5, 7)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = ours(xa), ref(xb)
    assert torch.allclose(ya, yb, rtol=1e-5, atol=1e-6)
    g = torch.randn_like(ya)
    ya.backward(g)
    yb.backward(g)
    assert torch.allclose(xa.grad, xb.grad, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ours.weight.grad, ref.weight.grad, rtol=1e-5, atol=1e-5)
    assert torch.allclose(ours.bias.grad, ref.bias.grad, rtol=1e-5, atol=1e-5)


def test_fused_loss_entry_points_reject_cpu_tensors():
    """no CPU path behind the fused losses either: they raise instead of computing elsewhere"""
    from scda_b200 import _lib
    from scda_b200.loss_ops import bce_sigmoid_rows, smooth_l1_masked_sum
    with pytest.raises(_lib.ScdaLibraryError):
        smooth_l1_masked_sum(torch.zeros(4), None, torch.zeros(4), 3.0)
    with pytest.raises(_lib.ScdaLibraryError):
        bce_sigmoid_rows(torch.zeros(2, 4), torch.ones(1, 4))


def test_smooth_l1_reference_formula_matches_oracle():
    """the tensor-op chain kept for CPU tensors = the oracle's restatement of the reference formula"""
    from oracle import host
    from scda_b200.models.faster_rcnn.faster_rcnn_adver_expansion_reweight_cluster import (
        _smooth_l1_masked, smooth_l1_loss_with_sigma)
    r = np.random.RandomState(0)
    p, t = r.randn(6, 36).astype(np.float32) * 0.3, r.randn(6, 36).astype(np.float32) * 0.3
    m = (r.rand(6, 36) > 0.5).astype(np.float32)
    want = host.smooth_l1_loss_with_sigma((p * m).astype(np.float64), t.astype(np.float64), 3.0)
    got = float(_smooth_l1_masked(torch.from_numpy(p), torch.from_numpy(m), torch.from_numpy(t)))
    assert abs(got - float(want)) <= 1e-5 * max(1.0, abs(float(want)))
    assert abs(float(smooth_l1_loss_with_sigma(torch.from_numpy(p * m), torch.from_numpy(t))) - got) <= 1e-6


def test_short_reduction_layers_are_not_split(monkeypatch):
    """conv5_x / RPN (16 pixel tiles, 4 x 4 x 9 = 144 (tap, tile) CTAs): one unsplit pass, so the kernel writes the
    gradient buffer itself; layers with too few tiles to fill the GPU, or with a long reduction, keep their split."""
    from scda_b200 import tc
    monkeypatch.setattr(tc, "WGRAD3", True)
    monkeypatch.setattr(tc, "SHORT_K_ONE_TAP", True)
    assert tc._wgrad_splits(1, 32, 64, 512, 512, 148) == 1
    assert tc._wgrad_splits(1, 32, 64, 64, 64, 148) > 1          # 9 CTAs per pass: split to fill the SMs
    assert tc._wgrad_splits(1, 64, 128, 512, 512, 148) > 1        # 64 pixel tiles: the three-tap split form
    monkeypatch.setattr(tc, "SHORT_K_ONE_TAP", False)
    assert tc._wgrad_splits(1, 32, 64, 512, 512, 148) > 1


def test_fresh_small_gradients_are_already_zero():
    """FlatGradBucket.zero() zeroes every gradient except the large direct-written ones and marks all direct ones
    fresh; a sink that finds a SMALL gradient fresh therefore clears nothing (tc_detector._clear_fresh), a large
    one is cleared by the sink itself."""
    from scda_b200.tc_detector import _clear_fresh, _take_fresh
    from scda_b200.utils.distributed_utils import DIRECT_SKIP_NUMEL, FlatGradBucket
    small = torch.nn.Parameter(torch.zeros(64))
    large = torch.nn.Parameter(torch.zeros(DIRECT_SKIP_NUMEL))
    plain = torch.nn.Parameter(torch.zeros(32))
    bucket = FlatGradBucket([small, large, plain])
    for p in (small, large):
        p._scda_direct_grad = True
    bucket.flat.fill_(3.0)
    bucket.zero()
    assert float(small.grad.abs().sum()) == 0 and float(plain.grad.abs().sum()) == 0
    assert float(large.grad.min()) == 3.0, "the large direct gradient is left for its writer"
    assert small._scda_grad_fresh and large._scda_grad_fresh and not getattr(plain, "_scda_grad_fresh", False)
    assert _take_fresh(small) and not _take_fresh(small)
    _clear_fresh(small)
    assert _take_fresh(large)
    _clear_fresh(large)
    assert float(large.grad.abs().sum()) == 0
    # a large gradient nobody wrote this step is zeroed by settle()
    bucket.flat.fill_(5.0)
    bucket.zero()
    bucket.settle()
    assert float(bucket.flat.abs().sum()) == 0
