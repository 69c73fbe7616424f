"""RoIAlign as an autograd op over libscda_b200.

Mirrors extensions/_roi_align/functions/roi_align.py:7-51:
`RoIAlignFunction(aligned_height, aligned_width, spatial_scale)(features, rois)`.
CUDA only, as in the reference (`raise NotImplementedError` on CPU, :30-31).
"""
import torch
from torch.autograd import Function

from ...._lib import check, load, require_cuda, stream_ptr


class _RoIAlignOp(Function):
    @staticmethod
    def forward(ctx, features, rois, aligned_height, aligned_width, spatial_scale):
        if not features.is_cuda:
            raise NotImplementedError
        require_cuda(features, rois)
        assert features.is_contiguous()
        assert rois.is_contiguous()
        assert features.dtype == torch.float32 and rois.dtype == torch.float32
        if rois.dim() != 2 or rois.size(1) != 5:
            raise ValueError("rois must be [R, 5] (batch, x1, y1, x2, y2)")
        batch_size, num_channels, data_height, data_width = features.size()
        num_rois = rois.size(0)
        output = features.new_empty(num_rois, num_channels, aligned_height, aligned_width)
        with torch.cuda.device(features.device):
            check(load().ROIAlignForwardLaucher(
                features.data_ptr(), spatial_scale, num_rois, data_height, data_width,
                num_channels, aligned_height, aligned_width, rois.data_ptr(), output.data_ptr(),
                stream_ptr(features.device)), "ROIAlignForwardLaucher")
        ctx.feature_size = features.size()
        ctx.align = (aligned_height, aligned_width, spatial_scale)
        ctx.save_for_backward(rois)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        aligned_height, aligned_width, spatial_scale = ctx.align
        batch_size, num_channels, data_height, data_width = ctx.feature_size
        assert grad_output.is_cuda
        grad_output = grad_output.contiguous()
        # the kernel accumulates into a zeroed buffer, as the reference's does (:40-41)
        grad_input = rois.new_zeros(batch_size, num_channels, data_height, data_width)
        with torch.cuda.device(grad_output.device):
            check(load().ROIAlignBackwardLaucher(
                grad_output.data_ptr(), spatial_scale, batch_size, rois.size(0), data_height,
                data_width, num_channels, aligned_height, aligned_width, rois.data_ptr(),
                grad_input.data_ptr(), stream_ptr(grad_output.device)), "ROIAlignBackwardLaucher")
        return grad_input, None, None, None, None


class RoIAlignFunction(object):
    def __init__(self, aligned_height, aligned_width, spatial_scale):
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def __call__(self, features, rois):
        return _RoIAlignOp.apply(features, rois, self.aligned_height, self.aligned_width,
                                 self.spatial_scale)

    forward = __call__
