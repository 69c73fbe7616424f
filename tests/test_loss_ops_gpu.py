"""Fused loss kernels (csrc/loss_ops.cu) through the C ABI against the oracle's restatement of the
reference formula (oracle/host.py: smooth_l1_loss_with_sigma, reference
models/faster_rcnn/faster_rcnn_adver_expansion_reweight_cluster.py:238-246) and against the tensor-op
chain the reference executes for the adversarial terms (torch.sigmoid + F.binary_cross_entropy,
tools/faster_rcnn_train_val.py:577-600).  fp32 sums in a different order: 1e-5 relative (north_star: 1e-4)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 60, 32, 64), (512, 36), (7,), (3, 1025)])
@pytest.mark.parametrize("with_mask", [True, False])
def test_smooth_l1_masked_sum(cuda_lib, shape, with_mask):
    import torch
    from oracle import host
    from scda_b200.loss_ops import smooth_l1_masked_sum
    from scda_b200.models.faster_rcnn.faster_rcnn_adver_expansion_reweight_cluster import smooth_l1_loss_with_sigma
    g = torch.Generator().manual_seed(len(shape) * 17 + shape[-1])
    pred = (torch.randn(shape, generator=g) * 0.3)
    target = (torch.randn(shape, generator=g) * 0.3)
    mask = (torch.rand(shape, generator=g) > 0.6).float() if with_mask else None
    p = pred.cuda().requires_grad_(True)
    loss = smooth_l1_masked_sum(p, mask.cuda() if with_mask else None, target.cuda(), 3.0)
    (loss * 0.37).backward()
    eff = pred * mask if with_mask else pred
    want = host.smooth_l1_loss_with_sigma(eff.numpy().astype(np.float64), target.numpy().astype(np.float64), 3.0)
    assert abs(float(loss) - float(want)) <= 1e-5 * max(1.0, abs(float(want)))
    # gradient against autograd through the reference's own chain of tensor ops
    q = pred.clone().double().requires_grad_(True)
    ref = smooth_l1_loss_with_sigma(q * mask.double() if with_mask else q, target.double(), 3.0)
    (ref * 0.37).backward()
    assert torch.allclose(p.grad.cpu().double(), q.grad, rtol=1e-5, atol=1e-7)
    # elements sitting exactly on the |d| = 1/sigma^2 switch take the linear branch on both sides
    d = torch.tensor([1.0 / 9.0, -1.0 / 9.0, 0.0, 0.05])
    z = torch.zeros(4)
    got = float(smooth_l1_masked_sum(d.cuda(), None, z.cuda(), 3.0))
    assert abs(got - float(smooth_l1_loss_with_sigma(d, z, 3.0))) < 1e-6


@pytest.mark.parametrize("K,M", [(4, 1024), (4, 512), (1, 1), (3, 1500)])
@pytest.mark.parametrize("label", ["row", "one", "zero", "const"])
def test_bce_sigmoid_rows(cuda_lib, K, M, label):
    import torch
    import torch.nn.functional as F
    from scda_b200.loss_ops import bce_sigmoid_rows
    g = torch.Generator().manual_seed(K * 131 + M)
    x = torch.randn(K, M, generator=g) * 4          # saturating logits included
    x[0, 0] = 40.0                                  # p == 1 in fp32: the -100 clamp and the 1e-12 floor
    if M > 1:
        x[-1, -1] = -120.0                          # p == 0
    if label == "row":
        y = 0.8 + 0.2 * torch.rand(1, M, generator=g)
    elif label == "const":
        y = torch.full((1,), 0.15)
    else:
        y = torch.ones(1, M) if label == "one" else torch.zeros(1, M)
    w = torch.rand(K, generator=g)
    xa = x.cuda().requires_grad_(True)
    out = bce_sigmoid_rows(xa, y.cuda())
    (out * w.cuda()).sum().backward()
    xr = x.clone().cuda().requires_grad_(True)       # the reference's chain, on the same device
    p = torch.sigmoid(xr)
    ref = F.binary_cross_entropy(p, y.cuda().expand_as(p) if y.numel() > 1 else y.cuda().expand(K, M),
                                 reduction='none').mean(dim=1)
    (ref * w.cuda()).sum().backward()
    assert out.shape == (K,)
    assert torch.allclose(out, ref.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(xa.grad, xr.grad, rtol=1e-4, atol=1e-8)
    # deterministic
    assert torch.equal(bce_sigmoid_rows(x.cuda(), y.cuda()), out.detach())


@pytest.mark.parametrize("M,C,ignore", [(30720, 2, -1), (512, 9, -100), (7, 3, -1), (1000, 32, -1)])
def test_softmax_ce_acc_matches_torch(cuda_lib, M, C, ignore):
    """fused cross entropy (+ ignore index) and top-1 accuracy against F.cross_entropy and the reference's
    accuracy() (faster_rcnn_adver_expansion_reweight_cluster.py:249-267), forward and backward"""
    import torch
    import torch.nn.functional as F
    from scda_b200.loss_ops import softmax_ce_acc
    from scda_b200.models.faster_rcnn.faster_rcnn_adver_expansion_reweight_cluster import accuracy
    g = torch.Generator(device="cuda").manual_seed(M + C)
    x = (torch.randn(M, C, device="cuda", generator=g) * 2).requires_grad_(True)
    t = torch.randint(0, C, (M,), device="cuda", generator=g)
    if ignore == -1:
        t = torch.where(torch.rand(M, device="cuda", generator=g) < 0.6, torch.full_like(t, -1), t)
    loss, acc = softmax_ce_acc(x, t, ignore_index=ignore)
    (loss * 1.7).backward()
    gx = x.grad.clone()
    x.grad = None
    ref = F.cross_entropy(x, t, ignore_index=ignore)
    (ref * 1.7).backward()
    assert abs(float(loss) - float(ref)) <= 2e-6 * max(1.0, abs(float(ref)))
    assert torch.allclose(gx, x.grad, rtol=1e-5, atol=1e-8)
    ref_acc = accuracy(x.detach(), t, ignore_index=ignore)[0]
    assert abs(float(acc) - float(ref_acc)) < 1e-3


def test_softmax_ce_all_ignored_is_nan(cuda_lib):
    import torch
    from scda_b200.loss_ops import softmax_ce_acc
    x = torch.randn(64, 2, device="cuda")
    loss, acc = softmax_ce_acc(x, torch.full((64,), -1, device="cuda"), ignore_index=-1)
    assert bool(torch.isnan(loss)) and float(acc) == 0.0


def test_rpn_fg_scores_matches_softmax(cuda_lib):
    import torch
    import torch.nn.functional as F
    from scda_b200.loss_ops import rpn_fg_scores
    g = torch.Generator(device="cuda").manual_seed(5)
    cls = torch.randn(2, 30, 32, 64, device="cuda", generator=g) * 3
    x = cls.permute(0, 2, 3, 1).contiguous()
    ref = F.softmax(x.view(-1, 2), dim=1).view(2, -1, 2)[..., 1]
    got = rpn_fg_scores(cls)
    assert got.shape == ref.shape and float((got - ref).abs().max()) < 1e-6
