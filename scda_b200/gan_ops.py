"""Fused channels-last operators of the reconstruction networks (csrc/norm_ops.cu).

`instance_norm_act(x, act)` = nn.InstanceNorm2d(affine=False, eps=1e-5) followed by nothing /
ReLU / LeakyReLU, on a channels_last fp32 CUDA tensor, forward and backward, without the
NCHW round trip torch's instance norm forces (models/faster_rcnn/common_net.py:59-80,
279-293 of the reference use the torch modules)."""
import logging
import os

import torch

from ._lib import check, load, require_cuda, stream_ptr

_ACT = {None: 0, "none": 0, "relu": 1, "leaky_relu": 2}

# CUDA-path calls that went to a LIBRARY (cuDNN / torch) instead of a hand-written kernel, by layer kind:
# counted, logged once per kind at WARNING, and reported by bench.py as `library_fallbacks`
LIBRARY_CALLS = {}
_log = logging.getLogger("scda_b200")


def note_library_call(kind, why=""):
    n = LIBRARY_CALLS.get(kind, 0)
    if n == 0:
        _log.warning("scda_b200: %s runs on a library kernel (cuDNN / torch), not a hand-written one%s",
                     kind, (": " + why) if why else "")
    LIBRARY_CALLS[kind] = n + 1


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
    def forward(ctx, x, out_bf16):
        require_cuda(x)
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        N, C, H, W = x.shape
        y = torch.empty((N, C, 2 * H, 2 * W), dtype=torch.bfloat16 if out_bf16 else x.dtype, device=x.device,
                        memory_format=torch.channels_last)
        with torch.cuda.device(x.device):
            check(load().scda_upsample_bilinear2x_nhwc(N, H, W, C, x.data_ptr(), y.data_ptr(),
                                                       1 if out_bf16 else 0, stream_ptr(x.device)),
                  "scda_upsample_bilinear2x_nhwc")
        ctx.shape = (N, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, C, H, W = ctx.shape
        if dy.dtype not in (torch.float32, torch.bfloat16):
            dy = dy.float()
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, C, H, W), dtype=torch.float32, device=dy.device, memory_format=torch.channels_last)
        with torch.cuda.device(dy.device):
            check(load().scda_upsample_bilinear2x_bwd_nhwc(N, H, W, C, dy.data_ptr(),
                                                           1 if dy.dtype == torch.bfloat16 else 0, dx.data_ptr(),
                                                           stream_ptr(dy.device)),
                  "scda_upsample_bilinear2x_bwd_nhwc")
        return dx, None


def upsample_bilinear2x(x, out_bf16=False):
    """F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True) on a channels_last
    fp32 CUDA tensor (C % 4 == 0, H, W >= 2); out_bf16 writes the result as bf16 (the operand
    dtype of the tensor-core convolution behind it; ignored in the fp32-parity mode)."""
    from . import tc
    return _Upsample2x.apply(x, bool(out_bf16) and not tc.x3())


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


# ---------------------------------------------------------------------------------------
# Two independent sub-networks side by side (decode_A | decode_B, model_A | model_B, the
# discriminator on reconstructions | on real crops): each is a chain of small kernels that
# fills a fraction of the 148 SMs, so the second one runs on a helper stream forked from and
# joined back into the current one.  Off unless the engine turns it on (PAIR_STREAMS).
PAIR_STREAMS = False
_HELPERS = {}


def _helper_stream(cur):
    key = (cur.device.index, cur.cuda_stream)
    st = _HELPERS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=cur.device, priority=cur.priority)
        _HELPERS[key] = st
    return st


# Weight / bias gradients of the decoder's convolution nodes on a helper stream: nothing in the backward reads
# them (they are reduced straight into the flat gradient buffer), so the data-gradient chain of phase 3 —
# InstanceNorm backward -> dgrad -> next node, the tail of the iteration's critical path — does not queue behind
# colsum + weight gradient + slab reduction of every node.  Set by the engine; joined by `join_wgrad_streams()`
# before the gradients are used (all-reduce / Adam).
WGRAD_SIDE = False
_WG_PENDING = []


def _wgrad_stream(cur):
    key = ("wg", cur.device.index, cur.cuda_stream)
    st = _HELPERS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=cur.device, priority=cur.priority)
        _HELPERS[key] = st
    return st


def run_wgrad_side(fn, *operands):
    """fn() -> tuple of gradients (None = written in place).  With WGRAD_SIDE it runs on the helper stream of
    the current stream; `operands`: the tensors it reads that the caller may release before it is done."""
    if not (WGRAD_SIDE and torch.cuda.is_available()):
        return fn()
    cur = torch.cuda.current_stream()
    side = _wgrad_stream(cur)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        out = fn()
    for t in operands:
        if t is not None and t.is_cuda:
            t.record_stream(side)
    if all(o is None for o in out):
        _WG_PENDING.append(side)              # joined by join_wgrad_streams()
    else:
        cur.wait_stream(side)                 # handed to autograd on this stream
    return out


def join_wgrad_streams():
    """the current stream waits for every helper stream that still carries weight-gradient work"""
    if not _WG_PENDING:
        return
    cur = torch.cuda.current_stream()
    seen = set()
    while _WG_PENDING:
        st = _WG_PENDING.pop()
        if st.cuda_stream not in seen:
            seen.add(st.cuda_stream)
            cur.wait_stream(st)


def _tensors(obj):
    if torch.is_tensor(obj):
        yield obj
    elif isinstance(obj, (tuple, list)):
        for o in obj:
            for t in _tensors(o):
                yield t


def run_pair(fa, fb):
    """(fa(), fb()); on CUDA with PAIR_STREAMS fb runs on a helper stream beside fa."""
    if not (PAIR_STREAMS and torch.cuda.is_available()):
        return fa(), fb()
    cur = torch.cuda.current_stream()
    side = _helper_stream(cur)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        b = fb()
    a = fa()
    cur.wait_stream(side)
    for t in _tensors(b):
        if t.is_cuda:
            t.record_stream(cur)        # allocated on the helper, consumed (and freed) on `cur`
    return a, b


# ---------------------------------------------------------------------------------------
# conv3x3 (+ bias) -> InstanceNorm -> activation as ONE autograd node on the tensor cores.
# The decoder's stride-1 3x3 convolutions with 64-multiple channel counts (6 residual-block
# convolutions + the first up-sampling convolution per decoder, ~80 % of the decoder's flops;
# models/faster_rcnn/common_net.py:59-80, 279-293 of the reference, cuDNN fp32 there) run on
# the same tcgen05 halo kernel as the backbone (csrc/conv_halo.cu): bf16 operands, fp32
# accumulation, fp32 convolution output (the InstanceNorm statistics are taken in fp32).
# Fusing the three layers into one node lets every tensor cross a kernel boundary in the dtype
# its consumer wants (IN output / IN input-gradient in bf16 for the next MMA) with no cast pass.
TC_GAN = os.environ.get("SCDA_GAN_TC", "1") != "0"
# 32 output channels (the decoder's last 3x3 layer) ride in a 64-wide tile with TMA zero fill; "0" leaves that
# layer on cuDNN (A/B knob)
_TC_MIN_OUT = 64 if os.environ.get("SCDA_GAN_TC32", "1") == "0" else 32


class _ConvINActTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act, slope, eps, out_bf16):
        from . import tc
        from .tc_detector import shadow3_of, shadow_of
        require_cuda(x)
        x3 = tc.x3()          # fp32-parity mode: fp32 activations, split operands (csrc/x3_ops.cu)
        out_bf16 = out_bf16 and not x3
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        if x3:
            xb = tc.split3(x.float().permute(0, 2, 3, 1)).permute(0, 3, 1, 2)   # [N,3C,H,W] view of NHWC
            w = shadow3_of(weight)[0]                                           # bf16 [O,3,3,3I]
        else:
            xb = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)       # keeps channels_last
            w = shadow_of(weight)                                               # bf16 [O,3,3,I]
        xn = xb.permute(0, 2, 3, 1)                                         # [N,H,W,C] contiguous view
        N, H, W, _ = xn.shape
        O = w.shape[0]
        c = tc.conv3x3_nhwc(xn, w, bias.detach() if bias is not None else None, out_dtype=torch.float32)
        y = torch.empty(N, H, W, O, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=x.device)
        mean = torch.empty(N, O, dtype=torch.float32, device=x.device)
        rstd = torch.empty(N, O, dtype=torch.float32, device=x.device)
        lib = load()
        wsb = lib.scda_instnorm_workspace_bytes(N, H * W, O)
        if wsb == 0:
            raise ValueError("conv_in_act_tc: unsupported shape %s" % (tuple(c.shape),))
        ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.scda_instnorm_act_fwd_nhwc(N, H * W, O, c.data_ptr(), y.data_ptr(), 1 if out_bf16 else 0,
                                                 mean.data_ptr(), rstd.data_ptr(), eps, act, slope, ws.data_ptr(),
                                                 wsb, stream_ptr(x.device)), "scda_instnorm_act_fwd_nhwc")
        ctx.save_for_backward(xb, c, mean, rstd)
        ctx.params = (weight, bias)
        ctx.cfg = (act, slope, x.dtype, x3)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        from . import tc
        from .tc_detector import (_sink_bias, _sink_bias_f32, _sink_conv_wgrad, _sink_conv_wgrad_x3, shadow3_of,
                                  shadow_of)
        xb, c, mean, rstd = ctx.saved_tensors
        weight, bias = ctx.params
        act, slope, x_dtype, x3 = ctx.cfg
        N, H, W, O = c.shape
        if dy.dtype not in (torch.float32, torch.bfloat16):
            dy = dy.float()
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dc = torch.empty(N, H, W, O, dtype=torch.float32 if x3 else torch.bfloat16, device=c.device)
        lib = load()
        wsb = lib.scda_instnorm_workspace_bytes(N, H * W, O) + 8 * N * O
        ws = torch.empty(wsb, dtype=torch.uint8, device=c.device)
        with torch.cuda.device(c.device):
            check(lib.scda_instnorm_act_bwd_nhwc(N, H * W, O, c.data_ptr(), dy.data_ptr(),
                                                 1 if dy.dtype == torch.bfloat16 else 0, mean.data_ptr(),
                                                 rstd.data_ptr(), dc.data_ptr(), 0 if x3 else 1, act, slope,
                                                 ws.data_ptr(), wsb, stream_ptr(c.device)),
                  "scda_instnorm_act_bwd_nhwc")
        xn = xb.permute(0, 2, 3, 1)
        dx = None
        if x3:
            gb = _sink_bias_f32(bias, dc.view(-1, O)) if bias is not None and ctx.needs_input_grad[2] else None
            ds = tc.split3(dc)
            gw = _sink_conv_wgrad_x3(weight, xn, ds) if ctx.needs_input_grad[1] else None
            if ctx.needs_input_grad[0]:
                dx = tc.conv3x3_dgrad_nhwc(ds, shadow3_of(weight)[1], out_dtype=torch.float32).permute(0, 3, 1, 2)
            return dx, gw, gb, None, None, None, None
        need = ctx.needs_input_grad
        # (both operands are released when this node returns, while the helper stream may still be reading them)
        gb, gw = run_wgrad_side(lambda: (_sink_bias(bias, dc.view(-1, O)) if bias is not None and need[2] else None,
                                         _sink_conv_wgrad(weight, xn, dc) if need[1] else None), dc, xb)
        if ctx.needs_input_grad[0]:
            dx = tc.conv3x3_dgrad_nhwc(dc, shadow_of(weight), out_dtype=x_dtype).permute(0, 3, 1, 2)
        return dx, gw, gb, None, None, None, None


def conv_in_act_tc_supported(x, conv):
    return (TC_GAN and x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16)
            and conv.kernel_size == (3, 3) and conv.stride == (1, 1) and conv.padding == (1, 1)
            and conv.dilation == (1, 1) and conv.groups == 1 and conv.padding_mode == 'zeros'
            and conv.in_channels % 64 == 0 and conv.out_channels % _TC_MIN_OUT == 0 and conv.out_channels <= 1024
            and 1024 % conv.out_channels == 0 and x.shape[3] % 8 == 0
            and conv.weight.dtype == torch.float32)


def conv_in_act_tc(x, conv, act=None, negative_slope=0.01, eps=1e-5, out_bf16=False):
    """act(InstanceNorm(conv(x))) for an nn.Conv2d(k=3, s=1, p=1); returns a channels_last tensor
    (fp32, or bf16 with out_bf16)."""
    return _ConvINActTC.apply(x, conv.weight, conv.bias, _ACT[act], float(negative_slope), float(eps),
                              bool(out_bf16))


# ---------------------------------------------------------------------------------------
# Decoder head: ConvTranspose2d(32 -> 3, k = 1) + Tanh as one streaming kernel each way
# (csrc/loss_ops.cu: scda_conv1x1_tanh_*).
class _Conv1x1Tanh(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        require_cuda(x)
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        N, Cin, H, W = x.shape
        Cout = weight.shape[1]
        w2 = weight.detach().reshape(Cin, Cout).contiguous()
        y = torch.empty(N, H, W, Cout, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(load().scda_conv1x1_tanh_fwd(N * H * W, Cin, Cout, x.data_ptr(), w2.data_ptr(),
                                               bias.data_ptr() if bias is not None else None, y.data_ptr(),
                                               stream_ptr(x.device)), "scda_conv1x1_tanh_fwd")
        ctx.save_for_backward(x, w2, y)
        ctx.has_bias = bias is not None
        ctx.wshape = tuple(weight.shape)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        x, w2, y = ctx.saved_tensors
        N, Cin, H, W = x.shape
        Cout = w2.shape[1]
        dy = dy.float()
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        lib = load()
        P = N * H * W
        wsb = lib.scda_conv1x1_tanh_workspace_bytes(P, Cin, Cout)
        ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
        dx = torch.empty(N, H, W, Cin, dtype=torch.float32, device=x.device) if ctx.needs_input_grad[0] else None
        dw = torch.empty(Cin, Cout, dtype=torch.float32, device=x.device)
        db = torch.empty(Cout, dtype=torch.float32, device=x.device) if ctx.has_bias else None
        with torch.cuda.device(x.device):
            check(lib.scda_conv1x1_tanh_bwd(P, Cin, Cout, x.data_ptr(), w2.data_ptr(), y.data_ptr(), dy.data_ptr(),
                                            dx.data_ptr() if dx is not None else None, dw.data_ptr(),
                                            db.data_ptr() if db is not None else None, ws.data_ptr(), wsb,
                                            stream_ptr(x.device)), "scda_conv1x1_tanh_bwd")
        return (dx.permute(0, 3, 1, 2) if dx is not None else None), dw.view(ctx.wshape), db


def conv1x1_tanh_supported(x, convt):
    return (TC_GAN and x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and convt.in_channels == 32
            and 1 <= convt.out_channels <= 4 and convt.kernel_size == (1, 1) and convt.stride == (1, 1)
            and convt.padding == (0, 0) and convt.output_padding == (0, 0) and convt.groups == 1
            and convt.weight.dtype == torch.float32)


def conv1x1_tanh(x, convt):
    """tanh(convt(x)) for an nn.ConvTranspose2d(32, <= 4, kernel_size=1); returns a channels_last tensor
    tagged `_scda_tanh_applied` so that the nn.Tanh module behind it passes it through."""
    y = _Conv1x1Tanh.apply(x, convt.weight, convt.bias)
    y._scda_tanh_applied = True
    return y
