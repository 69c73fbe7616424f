"""`_RoIPooling` module (extensions/_roi_pooling/modules/roi_pool.py:5-15), exported as
`extensions.RoIPool`; the VGG16 detector's pooling op
(models/faster_rcnn/vgg_adver_expansion_cluster.py:45,75)."""
from torch.nn.modules.module import Module

from ..functions.roi_pool import RoIPoolFunction


class _RoIPooling(Module):
    def __init__(self, pooled_height, pooled_width, spatial_scale):
        super(_RoIPooling, self).__init__()
        self.pooled_width = int(pooled_width)
        self.pooled_height = int(pooled_height)
        self.spatial_scale = float(spatial_scale)

    def forward(self, features, rois):
        assert rois.shape[1] == 5
        return RoIPoolFunction(self.pooled_height, self.pooled_width, self.spatial_scale)(
            features, rois)

    def extra_repr(self):
        return "pooled=(%d, %d), spatial_scale=%g" % (self.pooled_height, self.pooled_width,
                                                      self.spatial_scale)
