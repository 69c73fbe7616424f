"""RoIAlign / RoIAlignAvg / RoIAlignMax modules
(extensions/_roi_align/modules/roi_align.py:6-44): the Avg/Max forms align to
(H+1) x (W+1) and then pool 2x2 with stride 1."""
from torch.nn.functional import avg_pool2d, max_pool2d
from torch.nn.modules.module import Module

from ..functions.roi_align import RoIAlignFunction


class RoIAlign(Module):
    def __init__(self, aligned_height, aligned_width, spatial_scale):
        super(RoIAlign, self).__init__()
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def forward(self, features, rois):
        return RoIAlignFunction(self.aligned_height, self.aligned_width, self.spatial_scale)(
            features, rois)


class RoIAlignAvg(RoIAlign):
    def forward(self, features, rois):
        assert rois.shape[1] == 5
        x = RoIAlignFunction(self.aligned_height + 1, self.aligned_width + 1,
                             self.spatial_scale)(features, rois)
        return avg_pool2d(x, kernel_size=2, stride=1)


class RoIAlignMax(RoIAlign):
    def forward(self, features, rois):
        assert rois.shape[1] == 5
        x = RoIAlignFunction(self.aligned_height + 1, self.aligned_width + 1,
                             self.spatial_scale)(features, rois)
        return max_pool2d(x, kernel_size=2, stride=1)
