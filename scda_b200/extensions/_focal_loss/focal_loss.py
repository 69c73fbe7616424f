"""Sigmoid / softmax focal losses over libscda_b200.

Mirrors extensions/_focal_loss/focal_loss.py:7-142:
`SigmoidFocalLossFunction(gamma, alpha, num_classes)(preds, targets, weight_pos)`
with preds [M, num_classes] fp32, targets [M] int32 (label d+1 <-> column d for
the sigmoid form, -1 = ignore), weight_pos a 1-element tensor; returns a
1-element CUDA tensor.  The reference's `losses.sum()` (:46, :118) is fused
into the kernel.
"""
import torch
from torch.autograd import Function

from ..._lib import check, load, require_cuda, stream_ptr


def _check_inputs(preds, targets, num_classes):
    require_cuda(preds, targets)
    assert preds.size(0) == targets.size(0)
    assert preds.size(1) == num_classes
    assert preds.is_contiguous()
    assert targets.is_contiguous()
    assert preds.dtype == torch.float32 and targets.dtype == torch.int32


class _SigmoidFocalOp(Function):
    @staticmethod
    def forward(ctx, preds, targets, weight_pos, gamma, alpha, num_classes):
        _check_inputs(preds, targets, num_classes)
        n = preds.numel()
        out = preds.new_empty(1)
        with torch.cuda.device(preds.device):
            check(load().scda_sigmoid_focal_loss_sum(
                n, preds.data_ptr(), targets.data_ptr(), weight_pos, gamma, alpha, num_classes,
                None, out.data_ptr(), stream_ptr(preds.device)), "scda_sigmoid_focal_loss_sum")
        ctx.save_for_backward(preds, targets)
        ctx.cfg = (weight_pos, gamma, alpha, num_classes)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        preds, targets = ctx.saved_tensors
        weight_pos, gamma, alpha, num_classes = ctx.cfg
        grad_input = torch.empty_like(preds)
        with torch.cuda.device(preds.device):
            check(load().SigmoidFocalLossBackwardLaucher(
                preds.numel(), preds.data_ptr(), targets.data_ptr(), grad_input.data_ptr(),
                weight_pos, gamma, alpha, num_classes, stream_ptr(preds.device)),
                "SigmoidFocalLossBackwardLaucher")
        return grad_input * grad_output, None, None, None, None, None


class _SoftmaxFocalOp(Function):
    @staticmethod
    def forward(ctx, preds, targets, weight_pos, gamma, alpha, num_classes):
        _check_inputs(preds, targets, num_classes)
        n = preds.numel()
        out = preds.new_empty(1)
        priors = torch.empty_like(preds)
        with torch.cuda.device(preds.device):
            check(load().scda_softmax_focal_loss_sum(
                n, preds.data_ptr(), targets.data_ptr(), weight_pos, gamma, alpha, num_classes,
                None, priors.data_ptr(), out.data_ptr(), stream_ptr(preds.device)),
                "scda_softmax_focal_loss_sum")
        ctx.save_for_backward(preds, targets, priors)
        ctx.cfg = (weight_pos, gamma, alpha, num_classes)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        preds, targets, priors = ctx.saved_tensors
        weight_pos, gamma, alpha, num_classes = ctx.cfg
        grad_input = torch.empty_like(preds)
        with torch.cuda.device(preds.device):
            check(load().SoftmaxFocalLossBackwardLaucher(
                preds.numel(), preds.data_ptr(), targets.data_ptr(), grad_input.data_ptr(),
                weight_pos, gamma, alpha, num_classes, priors.data_ptr(), None,
                stream_ptr(preds.device)), "SoftmaxFocalLossBackwardLaucher")
        return grad_input * grad_output, None, None, None, None, None


class _FocalLossFunction(object):
    _op = None

    def __init__(self, gamma, alpha, num_classes):
        self.gamma = float(gamma)
        self.alpha = float(alpha)
        self.num_classes = int(num_classes)

    def __call__(self, preds, targets, weight_pos):
        # the reference reads weight_pos on the host (focal_loss.py:27)
        weight_pos = float(weight_pos[0]) if torch.is_tensor(weight_pos) else float(weight_pos)
        return self._op.apply(preds, targets, weight_pos, self.gamma, self.alpha,
                              self.num_classes)

    forward = __call__


class SigmoidFocalLossFunction(_FocalLossFunction):
    _op = _SigmoidFocalOp


class SoftmaxFocalLossFunction(_FocalLossFunction):
    _op = _SoftmaxFocalOp
