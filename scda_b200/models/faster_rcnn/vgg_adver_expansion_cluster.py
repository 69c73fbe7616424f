"""VGG16 Faster R-CNN (mirrors models/faster_rcnn/vgg_adver_expansion_cluster.py of the
reference): torchvision-style VGG `features` with the last max-pool dropped (stride 16),
`NaiveRpnHead`, RoIPool 7x7 at 1/16, fc6/fc7 `classifier`, `fc_rcnn_cls` / `fc_rcnn_loc`.
Module and parameter names are the reference's, so `vgg16-397923af.pth` and SCDA
checkpoints load by name (utils/load_helper.py:28-54)."""
import math

import torch.nn as nn

from ...extensions import RoIPool
from ..head import NaiveRpnHead
from .faster_rcnn_adver_expansion_reweight_cluster import FasterRCNN_AdEx

__all__ = ['VGG', 'vgg11', 'vgg11_bn', 'vgg13', 'vgg13_bn', 'vgg16', 'vgg16_bn', 'vgg19_bn', 'vgg19']

cfg = {
    'A': [64, 'M', 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M'],
    'B': [64, 64, 'M', 128, 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M'],
    'D': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M'],
    'E': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M',
          512, 512, 512, 512, 'M'],
}


class VGG(FasterRCNN_AdEx):
    def __init__(self, features, cfg):
        super(VGG, self).__init__(cfg['gan_model_flag'])
        self.features = features
        # drop the last pooling layer so that the feature stride is 2^4
        last = list(self.features._modules.keys())[-1]
        del self.features._modules[last]
        num_anchors = len(cfg['anchor_scales']) * len(cfg['anchor_ratios'])
        self.rpn_head = NaiveRpnHead(512, num_classes=2, num_anchors=num_anchors)
        self.roipooling = RoIPool(7, 7, 1.0 / cfg['anchor_stride'])
        self.classifier = nn.Sequential(
            nn.Linear(512 * 7 * 7, 4096), nn.ReLU(True), nn.Dropout(),
            nn.Linear(4096, 4096), nn.ReLU(True), nn.Dropout())
        self.fc_rcnn_cls = nn.Linear(4096, cfg['num_classes'])
        self.fc_rcnn_loc = nn.Linear(4096, cfg['num_classes'] * 4)
        self._initialize_weights()

    # The three stages run on the tcgen05 kernels (scda_b200/tc_detector.py): the feature
    # map between them is NHWC bf16.  `_fp32_graph = True` runs the same modules through
    # torch.nn in fp32 NCHW instead — the plain-PyTorch reference the numerics tests
    # compare the kernels with, not a production path.
    _fp32_graph = False

    def _tc(self):
        rt = self.__dict__.get('_tc_runtime')
        if rt is None:
            from ...tc_detector import TcDetector
            rt = TcDetector(self)
            self.__dict__['_tc_runtime'] = rt
        return rt

    def feature_extractor(self, x):
        if self._fp32_graph:
            return self.features(x)
        return self._tc().features(x)

    def rpn(self, x):
        if self._fp32_graph:
            return self.rpn_head(x)
        return self._tc().rpn(x)

    def rcnn(self, x, rois):
        assert rois.shape[1] == 5
        if not self._fp32_graph:
            return self._tc().rcnn(x, rois)
        x = self.roipooling(x, rois)          # [R, 512, 7, 7]
        x = x.view(x.size(0), -1)
        x_fea = self.classifier(x)            # [R, 4096]
        return x_fea, self.fc_rcnn_cls(x_fea), self.fc_rcnn_loc(x_fea)

    def _initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()


def make_layers(cfg, batch_norm=False):
    layers = []
    in_channels = 3
    for v in cfg:
        if v == 'M':
            layers += [nn.MaxPool2d(kernel_size=2, stride=2)]
        else:
            conv2d = nn.Conv2d(in_channels, v, kernel_size=3, padding=1)
            layers += [conv2d, nn.BatchNorm2d(v), nn.ReLU(inplace=True)] if batch_norm \
                else [conv2d, nn.ReLU(inplace=True)]
            in_channels = v
    return nn.Sequential(*layers)


def _build(key, batch_norm, pretrained, **kwargs):
    if pretrained:
        raise RuntimeError("no network here: load ImageNet weights with utils.load_helper "
                           "from a local vgg16-397923af.pth instead of pretrained=True")
    return VGG(make_layers(cfg[key], batch_norm=batch_norm), **kwargs)


def vgg11(pretrained=False, **kwargs):
    return _build('A', False, pretrained, **kwargs)


def vgg11_bn(pretrained=False, **kwargs):
    return _build('A', True, pretrained, **kwargs)


def vgg13(pretrained=False, **kwargs):
    return _build('B', False, pretrained, **kwargs)


def vgg13_bn(pretrained=False, **kwargs):
    return _build('B', True, pretrained, **kwargs)


def vgg16(pretrained=False, **kwargs):
    return _build('D', False, pretrained, **kwargs)


def vgg16_bn(pretrained=False, **kwargs):
    return _build('D', True, pretrained, **kwargs)


def vgg19(pretrained=False, **kwargs):
    return _build('E', False, pretrained, **kwargs)


def vgg19_bn(pretrained=False, **kwargs):
    return _build('E', True, pretrained, **kwargs)
