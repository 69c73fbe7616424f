"""Fused channels-last operators of the reconstruction networks (csrc/norm_ops.cu).

`instance_norm_act(x, act)` = nn.InstanceNorm2d(affine=False, eps=1e-5) followed by nothing /
ReLU / LeakyReLU, on a channels_last fp32 CUDA tensor, forward and backward, without the
NCHW round trip torch's instance norm forces (models/faster_rcnn/common_net.py:59-80,
279-293 of the reference use the torch modules)."""
import torch

from ._lib import check, load, require_cuda, stream_ptr

_ACT = {None: 0, "none": 0, "relu": 1, "leaky_relu": 2}


class _InstNormAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act, slope, eps):
        require_cuda(x)
        assert x.dim() == 4 and x.dtype == torch.float32
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        N, C, H, W = x.shape
        y = torch.empty_like(x)                       # keeps the channels_last strides
        assert y.is_contiguous(memory_format=torch.channels_last)
        mean = torch.empty(N, C, dtype=torch.float32, device=x.device)
        rstd = torch.empty(N, C, dtype=torch.float32, device=x.device)
        lib = load()
        wsb = lib.scda_instnorm_workspace_bytes(N, H * W, C)
        if wsb == 0:
            raise ValueError("instance_norm_act: unsupported shape %s" % (tuple(x.shape),))
        ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.scda_instnorm_act_fwd_nhwc_f32(N, H * W, C, x.data_ptr(), y.data_ptr(), mean.data_ptr(),
                                                     rstd.data_ptr(), eps, act, slope, ws.data_ptr(), wsb,
                                                     stream_ptr(x.device)), "scda_instnorm_act_fwd_nhwc_f32")
        ctx.save_for_backward(x, mean, rstd)
        ctx.cfg = (act, slope)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        act, slope = ctx.cfg
        N, C, H, W = x.shape
        if dy.dtype != torch.float32 or not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.float().contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        lib = load()
        wsb = lib.scda_instnorm_workspace_bytes(N, H * W, C) + 8 * N * C
        ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.scda_instnorm_act_bwd_nhwc_f32(N, H * W, C, x.data_ptr(), dy.data_ptr(), mean.data_ptr(),
                                                     rstd.data_ptr(), dx.data_ptr(), act, slope, ws.data_ptr(),
                                                     wsb, stream_ptr(x.device)), "scda_instnorm_act_bwd_nhwc_f32")
        return dx, None, None, None


def instance_norm_act(x, act=None, negative_slope=0.01, eps=1e-5):
    return _InstNormAct.apply(x, _ACT[act], float(negative_slope), float(eps))


def supported(x):
    """shapes the fused kernel takes (C a multiple of 4 that divides 1024)"""
    c = x.shape[1]
    return x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and c % 4 == 0 and 1024 % c == 0


class _Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        require_cuda(x)
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        N, C, H, W = x.shape
        y = torch.empty((N, C, 2 * H, 2 * W), dtype=x.dtype, device=x.device,
                        memory_format=torch.channels_last)
        with torch.cuda.device(x.device):
            check(load().scda_upsample_bilinear2x_nhwc_f32(N, H, W, C, x.data_ptr(), y.data_ptr(),
                                                           stream_ptr(x.device)),
                  "scda_upsample_bilinear2x_nhwc_f32")
        ctx.shape = (N, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, C, H, W = ctx.shape
        if dy.dtype != torch.float32 or not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.float().contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, C, H, W), dtype=torch.float32, device=dy.device, memory_format=torch.channels_last)
        with torch.cuda.device(dy.device):
            check(load().scda_upsample_bilinear2x_bwd_nhwc_f32(N, H, W, C, dy.data_ptr(), dx.data_ptr(),
                                                               stream_ptr(dy.device)),
                  "scda_upsample_bilinear2x_bwd_nhwc_f32")
        return dx


def upsample_bilinear2x(x):
    """F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True) on a channels_last
    fp32 CUDA tensor (C % 4 == 0, H, W >= 2)."""
    return _Upsample2x.apply(x)


def upsample_supported(x, scale_factor, mode):
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and scale_factor == 2
            and mode == 'bilinear' and x.shape[1] % 4 == 0 and x.shape[2] >= 2 and x.shape[3] >= 2)


class _ConvBiasCL(torch.autograd.Function):
    """nn.Conv2d(bias=True) on channels_last fp32 CUDA tensors: cuDNN for the convolution and its
    input / weight gradients, scda_colsum_f32 for the bias gradient (torch reduces a channels_last
    gradient over (N, H, W) at ~1.5 TB/s: 0.5 ms per iteration over the reconstruction networks,
    profiles/r1_stagekernels_h.txt)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding):
        y = torch.nn.functional.conv2d(x, weight, bias, stride, padding)
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, padding)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        stride, padding = ctx.cfg
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dx, dw, _ = torch.ops.aten.convolution_backward(
            dy, x, weight, None, list(stride), list(padding), [1, 1], False, [0, 0], 1,
            [need_x, need_w, False])
        db = None
        if need_b:
            N, C, H, W = dy.shape
            db = torch.zeros(C, dtype=torch.float32, device=dy.device)
            with torch.cuda.device(dy.device):
                check(load().scda_colsum_f32(N * H * W, C, dy.data_ptr(), db.data_ptr(), stream_ptr(dy.device)),
                      "scda_colsum_f32")
        return dx, dw, db, None, None


def conv_bias_supported(x, conv):
    c = conv.out_channels
    g = c // 4
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and conv.bias is not None
            and conv.groups == 1 and conv.dilation == (1, 1) and conv.padding_mode == 'zeros'
            and isinstance(conv.padding, tuple)
            and c % 4 == 0 and g <= 256 and (g & (g - 1)) == 0 and torch.is_grad_enabled())


def conv2d_bias_cl(x, conv):
    """conv(x) for an nn.Conv2d with bias; x any layout (made channels_last)."""
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    return _ConvBiasCL.apply(x, conv.weight, conv.bias, conv.stride, conv.padding)
