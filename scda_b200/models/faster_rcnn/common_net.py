"""Building blocks of the reconstruction / discriminator networks (the subset of
models/faster_rcnn/common_net.py the SCDA driver instantiates: :15-19 gaussian init,
:59-80 INSResBlock, :107-129 LinUnsRes_cluster, :160-169 Interpolate, :205-245
ResDis_cluster, :251-261 LeakyReLUConv2d, :279-293 LeakyReLUConvTranspose2d_2).
Module / parameter names match the reference so state dicts are interchangeable."""
import torch
import torch.nn as nn

from ...gan_ops import (conv1x1_tanh, conv1x1_tanh_supported, conv2d_bias_cl, conv_bias_supported, conv_in_act_tc,
                        conv_in_act_tc_supported, instance_norm_act, note_library_call, supported as _fused_ok,
                        upsample_bilinear2x, upsample_supported)


class Conv2dCL(nn.Conv2d):
    """nn.Conv2d (same parameters, same state dict); on CUDA the bias gradient is computed by
    the library's channels-last column-sum kernel instead of torch's strided reduction."""

    def forward(self, x):
        if x.is_cuda:
            note_library_call("Conv2dCL %dx%d s%d" % (self.kernel_size + self.stride[:1]), "cuDNN convolution")
        if conv_bias_supported(x, self):
            return conv2d_bias_cl(x, self)
        return super(Conv2dCL, self).forward(x)


def gaussian_weights_init(m):
    name = m.__class__.__name__
    if name.find('Conv') == 0:
        m.weight.data.normal_(0.0, 0.02)


class INSResBlock(nn.Module):
    """conv3x3 - IN - ReLU - conv3x3 - IN (- Dropout) + identity."""

    def __init__(self, inplanes, planes, stride=1, dropout=0.0):
        super(INSResBlock, self).__init__()
        model = [Conv2dCL(inplanes, planes, kernel_size=3, stride=stride, padding=1),
                 nn.InstanceNorm2d(planes),
                 nn.ReLU(inplace=True),
                 Conv2dCL(planes, planes, kernel_size=3, stride=1, padding=1),
                 nn.InstanceNorm2d(planes)]
        if dropout > 0:
            model += [nn.Dropout(p=dropout)]
        self.model = nn.Sequential(*model)
        self.model.apply(gaussian_weights_init)

    def forward(self, x):
        m = self.model
        if conv_in_act_tc_supported(x, m[0]) and conv_in_act_tc_supported(x, m[3]):
            # conv + InstanceNorm (+ ReLU) as one node on the tensor cores; the block-internal
            # activation travels as bf16 (gan_ops._ConvINActTC)
            out = conv_in_act_tc(x, m[0], "relu", eps=m[1].eps, out_bf16=True)
            out = conv_in_act_tc(out, m[3], None, eps=m[4].eps)
            if len(m) > 5:
                out = m[5](out)
            return out + x
        if x.is_cuda:
            note_library_call("INSResBlock", "shape outside conv_in_act_tc_supported")
        if _fused_ok(x):
            # same layers, InstanceNorm + ReLU as one channels-last kernel (csrc/norm_ops.cu)
            out = instance_norm_act(m[0](x), "relu", eps=m[1].eps)
            out = instance_norm_act(m[3](out), None, eps=m[4].eps)
            if len(m) > 5:
                out = m[5](out)
            return out + x
        out = self.model(x)
        out += x
        return out


class LinUnsRes_cluster(nn.Module):
    """[cluster_num, threshold, 4096] -> view [cluster_num, channel, w, h]; no parameters."""

    def __init__(self, channel=128, w=64, h=64, cluster_num=4):
        super(LinUnsRes_cluster, self).__init__()
        self.channel, self.w, self.h, self.cluster_num = channel, w, h, cluster_num

    def forward(self, x):
        x = x.view(self.cluster_num, self.channel, self.w, self.h)
        # one transpose here keeps every cuDNN tensor-core convolution behind it in its
        # native NHWC layout (cuDNN otherwise transposes around each call)
        return x.contiguous(memory_format=torch.channels_last) if x.is_cuda else x


class Interpolate(nn.Module):
    def __init__(self, scale_factor, mode):
        super(Interpolate, self).__init__()
        self.scale_factor, self.mode = scale_factor, mode

    def forward(self, x):
        if upsample_supported(x, self.scale_factor, self.mode):
            return upsample_bilinear2x(x)          # channels-last kernel, csrc/norm_ops.cu
        if x.is_cuda:
            note_library_call("Interpolate", "scale / mode / shape outside upsample_supported")
        return nn.functional.interpolate(x, scale_factor=self.scale_factor, mode=self.mode,
                                         align_corners=True)


class ResDis_cluster(nn.Module):
    """Feature-level (patch) discriminator trunk: three stride-2 3x3 convs (BN + LeakyReLU
    after the first two), global average pool -> [cluster_num, 4 * n_in]."""

    def __init__(self, n_in=128, n_out=256, kernel_size=3, stride=2, padding=1, w=64, h=64,
                 cluster_num=4):
        super(ResDis_cluster, self).__init__()
        self.w, self.h, self.cluster_num, self.channel = w, h, cluster_num, n_in
        model = [nn.Conv2d(n_in, n_out, kernel_size, stride, padding, bias=False),
                 nn.BatchNorm2d(num_features=n_out),
                 nn.LeakyReLU(inplace=True),
                 nn.Conv2d(n_in * 2, n_out * 2, kernel_size, stride, padding, bias=False),
                 nn.BatchNorm2d(num_features=n_out * 2),
                 nn.LeakyReLU(inplace=True),
                 nn.Conv2d(n_out * 2, n_out * 2, kernel_size, stride, padding, bias=False)]
        self.model = nn.Sequential(*model)
        self.model.apply(gaussian_weights_init)

    def forward(self, x1):
        x1 = x1.view(self.cluster_num, self.channel, self.w, self.h)
        if x1.is_cuda:
            from ... import disc_ops
            if disc_ops.patch_dis_supported(self.model, x1):
                # the whole trunk as one autograd node on hand-written kernels (scda_b200/disc_ops.py)
                return torch.squeeze(disc_ops.patch_dis(self.model, x1))
            x1 = x1.contiguous(memory_format=torch.channels_last)
            note_library_call("ResDis_cluster", "stride-2 convolutions + BatchNorm on cuDNN "
                              "(fp32-parity mode, eval mode or an unsupported shape)")
        out = self.model(x1)
        out = nn.functional.avg_pool2d(out, out.size()[2:])
        return torch.squeeze(out)


class ConvTranspose1x1(nn.ConvTranspose2d):
    """nn.ConvTranspose2d(n_in, n_out, kernel_size=1, stride=1, padding=0) — same parameters
    ([n_in, n_out, 1, 1] weight, state-dict compatible) — evaluated as the equivalent 1x1
    convolution with the weight's first two axes swapped.  cuDNN's transposed-convolution
    weight gradient for the decoder's 32 -> 3 layer over 4 x 256 x 256 pixels takes 750 us
    per call (profiles/r1_launches_b_*: cutlass_80 s1688gemm tn_align1); the convolution
    form takes the regular wgrad path."""

    fuse_tanh = False      # set by the decoder when an nn.Tanh follows: one kernel computes both

    def forward(self, x, output_size=None):
        assert self.kernel_size == (1, 1) and self.stride == (1, 1) and self.padding == (0, 0)
        if self.fuse_tanh and conv1x1_tanh_supported(x, self):
            return conv1x1_tanh(x, self)         # tagged: the TanhAfterHead behind it passes it through
        if x.is_cuda:
            note_library_call("ConvTranspose1x1", "outside conv1x1_tanh_supported")
        return nn.functional.conv2d(x, self.weight.permute(1, 0, 2, 3), self.bias)


class TanhAfterHead(nn.Tanh):
    """nn.Tanh, except that a tensor the fused decoder head already passed through tanh is returned as is."""

    def forward(self, x):
        if getattr(x, "_scda_tanh_applied", False):
            return x
        return super(TanhAfterHead, self).forward(x)


class LeakyReLUConv2d(nn.Module):
    def __init__(self, n_in, n_out, kernel_size, stride, padding=0):
        super(LeakyReLUConv2d, self).__init__()
        self.model = nn.Sequential(
            Conv2dCL(n_in, n_out, kernel_size=kernel_size, stride=stride, padding=padding, bias=True),
            nn.LeakyReLU(inplace=True))
        self.model.apply(gaussian_weights_init)

    def forward(self, x):
        m = self.model
        if len(m) == 4 and isinstance(m[2], nn.InstanceNorm2d) and x.is_cuda:
            out = m[1](m[0](x))
            if _fused_ok(out):
                return instance_norm_act(out, "leaky_relu", m[3].negative_slope, m[2].eps)
            note_library_call("LeakyReLUConv2d InstanceNorm", "shape outside the fused kernel")
            return m[3](m[2](out))
        return self.model(x)


class LeakyReLUConvTranspose2d_2(nn.Module):
    """Despite the name: bilinear x2 upsample - conv - IN - LeakyReLU."""

    def __init__(self, n_in, n_out, kernel_size, stride, padding=0, output_padding=0):
        super(LeakyReLUConvTranspose2d_2, self).__init__()
        self.model = nn.Sequential(
            Interpolate(scale_factor=2, mode='bilinear'),
            Conv2dCL(in_channels=n_in, out_channels=n_out, kernel_size=kernel_size,
                      padding=padding, stride=1, bias=True),
            nn.InstanceNorm2d(num_features=n_out),
            nn.LeakyReLU(inplace=True))
        self.model.apply(gaussian_weights_init)

    def forward(self, x):
        m = self.model
        if len(m) == 4 and isinstance(m[2], nn.InstanceNorm2d) and x.is_cuda:
            if conv_in_act_tc_supported(x, m[1]) and upsample_supported(x, m[0].scale_factor, m[0].mode):
                up = upsample_bilinear2x(x, out_bf16=True)        # written as the MMA operand dtype
                return conv_in_act_tc(up, m[1], "leaky_relu", m[3].negative_slope, m[2].eps)
            note_library_call("LeakyReLUConvTranspose2d_2", "shape outside conv_in_act_tc_supported")
            out = m[1](m[0](x))
            if _fused_ok(out):
                return instance_norm_act(out, "leaky_relu", m[3].negative_slope, m[2].eps)
            return m[3](m[2](out))
        return self.model(x)
