"""RoI max pooling as an autograd op over libscda_b200.

Mirrors extensions/_roi_pooling/functions/roi_pool.py:6-42 of the reference:
`RoIPoolFunction(pooled_height, pooled_width, spatial_scale)(features, rois)`.
The reference class is a legacy instance-style autograd.Function (removed in
torch >= 1.3); the name and call signature are kept, the mechanism is a static
Function underneath.
"""
import torch
from torch.autograd import Function

from ...._lib import check, load, require_cuda, stream_ptr


class _RoIPoolOp(Function):
    @staticmethod
    def forward(ctx, features, rois, pooled_height, pooled_width, spatial_scale):
        require_cuda(features, rois)
        # same preconditions the reference asserts (roi_pool.py:25-26)
        assert features.is_contiguous()
        assert rois.is_contiguous()
        assert features.dtype == torch.float32 and rois.dtype == torch.float32
        if rois.dim() != 2 or rois.size(1) != 5:
            # roi_pooling_cuda.c:20-23 returns 0 here and the reference's Python ignores it
            raise ValueError("rois must be [R, 5] (batch, x1, y1, x2, y2)")
        batch_size, num_channels, data_height, data_width = features.size()
        num_rois = rois.size(0)
        output = features.new_empty(num_rois, num_channels, pooled_height, pooled_width)
        argmax = torch.empty(num_rois, num_channels, pooled_height, pooled_width,
                             dtype=torch.int32, device=features.device)
        with torch.cuda.device(features.device):
            check(load().ROIPoolForwardLaucher(
                features.data_ptr(), spatial_scale, num_rois, data_height, data_width,
                num_channels, pooled_height, pooled_width, rois.data_ptr(), output.data_ptr(),
                argmax.data_ptr(), stream_ptr(features.device)), "ROIPoolForwardLaucher")
        ctx.feature_size = features.size()
        ctx.pool = (pooled_height, pooled_width, spatial_scale)
        ctx.save_for_backward(rois, argmax)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        rois, argmax = ctx.saved_tensors
        pooled_height, pooled_width, spatial_scale = ctx.pool
        batch_size, num_channels, data_height, data_width = ctx.feature_size
        assert grad_output.is_cuda
        grad_output = grad_output.contiguous()
        grad_input = grad_output.new_empty(batch_size, num_channels, data_height, data_width)
        with torch.cuda.device(grad_output.device):
            check(load().ROIPoolBackwardLaucher(
                grad_output.data_ptr(), spatial_scale, batch_size, rois.size(0), data_height,
                data_width, num_channels, pooled_height, pooled_width, rois.data_ptr(),
                grad_input.data_ptr(), argmax.data_ptr(), stream_ptr(grad_output.device)),
                "ROIPoolBackwardLaucher")
        return grad_input, None, None, None, None


class RoIPoolFunction(object):
    def __init__(self, pooled_height, pooled_width, spatial_scale):
        self.pooled_width = int(pooled_width)
        self.pooled_height = int(pooled_height)
        self.spatial_scale = float(spatial_scale)

    def __call__(self, features, rois):
        return _RoIPoolOp.apply(features, rois, self.pooled_height, self.pooled_width,
                                self.spatial_scale)

    forward = __call__
