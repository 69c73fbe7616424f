"""RPN head (mirrors models/head.py:3-32 of the reference): 3x3 conv + ReLU, then two 1x1
convs for objectness (num_anchors * num_classes) and box deltas (num_anchors * 4).
Parameter names match the reference so checkpoints load unchanged."""
import torch.nn as nn


class NaiveRpnHead(nn.Module):
    def __init__(self, inplanes, num_classes, num_anchors):
        super(NaiveRpnHead, self).__init__()
        self.num_anchors, self.num_classes = num_anchors, num_classes
        self.conv3x3 = nn.Conv2d(inplanes, 512, kernel_size=3, stride=1, padding=1)
        self.relu3x3 = nn.ReLU(inplace=True)
        self.conv_cls = nn.Conv2d(512, num_anchors * num_classes, kernel_size=1, stride=1)
        self.conv_loc = nn.Conv2d(512, num_anchors * 4, kernel_size=1, stride=1)

    def forward(self, x):
        '''
        x: [B, inplanes, h, w] -> pred_cls [B, num_anchors*num_classes, h, w],
                                   pred_loc [B, num_anchors*4, h, w]
        '''
        x = self.relu3x3(self.conv3x3(x))
        return self.conv_cls(x), self.conv_loc(x)
