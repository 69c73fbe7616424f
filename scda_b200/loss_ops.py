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


class _SoftmaxCEAcc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, ignore_index):
        require_cuda(logits, targets)
        assert logits.dim() == 2 and logits.dtype == torch.float32 and logits.stride(1) == 1
        x = logits
        t = targets.contiguous()
        assert t.dtype == torch.int64 and t.numel() == x.shape[0]
        M, C = x.shape
        lib = load()
        wsb = lib.scda_softmax_ce_workspace_bytes(M)
        ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
        out3 = torch.empty(3, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.scda_softmax_ce_acc_fwd(M, C, x.data_ptr(), x.stride(0), t.data_ptr(), int(ignore_index),
                                              out3.data_ptr(), ws.data_ptr(), wsb, stream_ptr(x.device)),
                  "scda_softmax_ce_acc_fwd")
        ctx.save_for_backward(x, t, out3)
        ctx.ignore_index = int(ignore_index)
        acc = out3[1:2]
        ctx.mark_non_differentiable(acc)
        return out3[0], acc

    @staticmethod
    def backward(ctx, g, _g_acc):
        x, t, out3 = ctx.saved_tensors
        M, C = x.shape
        g = g.contiguous().float().reshape(1)
        dx = torch.empty(M, C, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(load().scda_softmax_ce_bwd(M, C, x.data_ptr(), x.stride(0), t.data_ptr(), ctx.ignore_index,
                                             out3.data_ptr(), g.data_ptr(), dx.data_ptr(), C,
                                             stream_ptr(x.device)), "scda_softmax_ce_bwd")
        return dx, None, None


def softmax_ce_acc(logits, targets, ignore_index=-100):
    """(F.cross_entropy(logits, targets, ignore_index=ignore_index), top-1 accuracy in percent as a
    1-element tensor): one pass over logits [M, C <= 32] fp32, int64 targets (csrc/loss_ops.cu)."""
    return _SoftmaxCEAcc.apply(logits, targets, int(ignore_index))


def rpn_fg_scores(cls_nchw):
    """[B, 2A, H, W] RPN class map -> [B, H*W*A] foreground probabilities in anchor order."""
    require_cuda(cls_nchw)
    x = cls_nchw.detach().contiguous().float()
    B, A2, H, W = x.shape
    out = torch.empty(B, H * W * (A2 // 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(load().scda_rpn_fg_scores(B, A2 // 2, H, W, x.data_ptr(), out.data_ptr(), stream_ptr(x.device)),
              "scda_rpn_fg_scores")
    return out
