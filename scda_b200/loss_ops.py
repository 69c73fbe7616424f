"""Fused loss operators (csrc/loss_ops.cu) behind torch.autograd:

`smooth_l1_masked_sum(pred, mask, target, sigma)` = the reference's
`smooth_l1_loss_with_sigma(pred * mask, target, sigma)`
(models/faster_rcnn/faster_rcnn_adver_expansion_reweight_cluster.py:238-246), one kernel each way.

`bce_sigmoid_rows(logits, label)` = `F.binary_cross_entropy(torch.sigmoid(logits[k:k+1]), label)` for
every row k (the per-cluster adversarial terms of tools/faster_rcnn_train_val.py:577-600, 655-680,
716-732) -> a [K] vector, one kernel each way instead of sigmoid + BCE + mean (+ their backwards)."""
import torch

from ._lib import check, load, require_cuda, stream_ptr


class _SmoothL1MaskedSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, mask, target, sigma):
        require_cuda(pred, target)
        pred_c, target_c = pred.contiguous(), target.contiguous()
        mask_c = mask.contiguous() if mask is not None else None
        assert pred_c.dtype == torch.float32 and target_c.dtype == torch.float32 and pred_c.shape == target_c.shape
        assert mask_c is None or (mask_c.dtype == torch.float32 and mask_c.shape == pred_c.shape)
        out = torch.empty((), dtype=torch.float32, device=pred.device)
        with torch.cuda.device(pred.device):
            check(load().scda_smooth_l1_sigma_sum_fwd(pred_c.numel(), pred_c.data_ptr(),
                                                      mask_c.data_ptr() if mask_c is not None else None,
                                                      target_c.data_ptr(), float(sigma), out.data_ptr(),
                                                      stream_ptr(pred.device)), "scda_smooth_l1_sigma_sum_fwd")
        ctx.save_for_backward(pred_c, target_c) if mask_c is None else ctx.save_for_backward(pred_c, target_c, mask_c)
        ctx.sigma = float(sigma)
        return out

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        pred, target = saved[0], saved[1]
        mask = saved[2] if len(saved) > 2 else None
        g = g.contiguous().float()
        gp = torch.empty_like(pred)
        with torch.cuda.device(pred.device):
            check(load().scda_smooth_l1_sigma_sum_bwd(pred.numel(), pred.data_ptr(),
                                                      mask.data_ptr() if mask is not None else None,
                                                      target.data_ptr(), ctx.sigma, g.data_ptr(), gp.data_ptr(),
                                                      stream_ptr(pred.device)), "scda_smooth_l1_sigma_sum_bwd")
        return gp, None, None, None


def smooth_l1_masked_sum(pred, mask, target, sigma=3.0):
    """sum of the sigma-smooth-L1 of (pred * mask - target); mask may be None.  CUDA fp32 tensors."""
    return _SmoothL1MaskedSum.apply(pred, mask, target, float(sigma))


class _BceSigmoidRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, label):
        require_cuda(logits, label)
        x = logits.contiguous()
        assert x.dim() == 2 and x.dtype == torch.float32
        K, M = x.shape
        y = label.detach().reshape(-1).contiguous().float()
        assert y.numel() in (1, M), "label: one row of M values or one constant"
        stride = 1 if y.numel() == M and M > 1 else (1 if M == 1 else 0)
        out = torch.empty(K, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(load().scda_bce_sigmoid_rows_fwd(K, M, x.data_ptr(), y.data_ptr(), stride, out.data_ptr(),
                                                   stream_ptr(x.device)), "scda_bce_sigmoid_rows_fwd")
        ctx.save_for_backward(x, y)
        ctx.stride = stride
        return out

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        K, M = x.shape
        g = g.contiguous().float()
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            check(load().scda_bce_sigmoid_rows_bwd(K, M, x.data_ptr(), y.data_ptr(), ctx.stride, g.data_ptr(),
                                                   gx.data_ptr(), stream_ptr(x.device)), "scda_bce_sigmoid_rows_bwd")
        return gx, None


def bce_sigmoid_rows(logits, label):
    """[K] vector: mean over M of BCE(sigmoid(logits[k]), label) — label [1, M] / [M] or a 1-element tensor."""
    return _BceSigmoidRows.apply(logits, label)
