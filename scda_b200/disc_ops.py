"""The two discriminators on hand-written kernels (csrc/disc_ops.cu, csrc/conv_halo.cu kS2,
csrc/gemm_tc.cu scda_conv3x3_s2_*), each network ONE autograd node whose backward walks the layers
by hand — the LeakyReLU / BatchNorm gradients ride in the kernels that produce the next MMA operand,
weight gradients go straight into the optimiser's flat gradient buffer.

Reference: `GAN_dis_AE` (image level: three LeakyReLUConv2d(k 3, s 2) + a 1x1 head per domain) and
`GAN_dis_AE_patch` / `ResDis_cluster` (feature level: conv s2 - BN - LReLU - conv s2 - BN - LReLU -
conv s2 - global average pool), models/faster_rcnn/faster_rcnn_adver_expansion_reweight_cluster.py:270-333,
common_net.py:205-261 — cuDNN fp32 + ~10 elementwise launches per layer there.

bf16 mode only (operands bf16, fp32 accumulation); in the fp32-parity mode (tc.set_precision('bf16x3'))
the modules fall back to cuDNN fp32 and say so (gan_ops.note_library_call)."""
import torch

from . import tc
from ._lib import check, load, require_cuda, stream_ptr
from .tc_detector import _clear_fresh, _direct, _is_krsc, _take_fresh, shadow_of

LEAKY, OUT_F32, MASK_POS, MASK_LEAKY = 128, 2, 4, 256

# `loss.backward(inputs=other_network.params)` still reports needs_input_grad = True for the discriminator's
# own weights (the flag is static), so the decoder's update (phase 3) would compute — and write into the
# discriminator's gradient buffer — weight gradients nobody uses.  The engine sets this around that backward:
# only the input gradient is produced.
FREEZE_PARAMS = False


# The mirror case: phase 1 differentiates the image discriminator's loss w.r.t. the discriminator's parameters
# only (`backward(inputs=dis.params)`), but the reconstructions it was applied to require grad (they come out of
# the decoder), so needs_input_grad[0] is True and the first layer's INPUT gradient (0.08-0.1 ms per decoder,
# l1_bwd_x_kernel) would be computed and dropped.  The engine sets this around that backward.
FREEZE_INPUT = False


class frozen_inputs(object):
    def __enter__(self):
        global FREEZE_INPUT
        self.prev, FREEZE_INPUT = FREEZE_INPUT, True

    def __exit__(self, *exc):
        global FREEZE_INPUT
        FREEZE_INPUT = self.prev


class frozen_params(object):
    def __enter__(self):
        global FREEZE_PARAMS
        self.prev, FREEZE_PARAMS = FREEZE_PARAMS, True

    def __exit__(self, *exc):
        global FREEZE_PARAMS
        FREEZE_PARAMS = self.prev


def _ptr(t):
    return t.data_ptr() if t is not None else None


def s2_weights(p):
    """bf16 [Cout][3][3][4C] phase-decomposed layout of conv parameter p ([Cout, C, 3, 3]), cached on the
    parameter and rebuilt when it changed"""
    stamp = (p._version, getattr(p, "_scda_epoch", 0))
    ent = getattr(p, "_scda_s2w", None)
    if ent is not None and ent[0] == stamp:
        return ent[1]
    sh = shadow_of(p)                                  # bf16 [O,3,3,C]
    O, C = sh.shape[0], sh.shape[3]
    wd = ent[1] if ent is not None else torch.empty(O, 3, 3, 4 * C, dtype=torch.bfloat16, device=p.device)
    with torch.cuda.device(p.device):
        check(load().scda_conv_s2_weights(O, C, sh.data_ptr(), wd.data_ptr(), stream_ptr(p.device)),
              "scda_conv_s2_weights")
    p._scda_s2w = (stamp, wd)
    return wd


def conv_s2(x, wd, bias, flags, slope=0.01, out_dtype=torch.bfloat16):
    """x bf16 NHWC [N, 2Ho, 2Wo, C] -> [N, Ho, Wo, Cout] (3x3, stride 2, padding 1)"""
    N, H, W, C = x.shape
    Cout = wd.shape[0]
    y = torch.empty(N, H // 2, W // 2, Cout, dtype=out_dtype, device=x.device)
    if out_dtype == torch.float32:
        flags |= OUT_F32
    with torch.cuda.device(x.device):
        check(load().scda_conv3x3_s2_bf16_nhwc(N, H // 2, W // 2, C, Cout, x.data_ptr(), wd.data_ptr(), _ptr(bias),
                                               y.data_ptr(), flags, float(slope), stream_ptr(x.device)),
              "scda_conv3x3_s2_bf16_nhwc")
    return y


def conv_s2_dgrad(dy, wd, C, mask_src=None, slope=0.01):
    """dy bf16 [N, Ho, Wo, Cout] -> dx bf16 [N, 2Ho, 2Wo, C]; with mask_src (= the conv's input, a LeakyReLU
    output) the gradient also passes through that LeakyReLU"""
    N, Ho, Wo, Cout = dy.shape
    dx = torch.empty(N, 2 * Ho, 2 * Wo, C, dtype=torch.bfloat16, device=dy.device)
    flags = (MASK_POS | MASK_LEAKY) if mask_src is not None else 0
    with torch.cuda.device(dy.device):
        check(load().scda_conv3x3_s2_dgrad_bf16_nhwc(N, Ho, Wo, C, Cout, dy.data_ptr(), wd.data_ptr(), dx.data_ptr(),
                                                     flags, _ptr(mask_src), float(slope), stream_ptr(dy.device)),
              "scda_conv3x3_s2_dgrad_bf16_nhwc")
    return dx


def conv_s2_wgrad(x, dy, out=None, accumulate=False):
    """dW fp32 [Cout, 3, 3, C] from x bf16 [N, 2Ho, 2Wo, C] and dy bf16 [N, Ho, Wo, Cout]"""
    N, H, W, C = x.shape
    Ho, Wo, Cout = dy.shape[1], dy.shape[2], dy.shape[3]
    tw = 16 if Wo % 16 == 0 else 8
    tiles = N * (-(-Ho // (128 // tw))) * (Wo // tw)          # 128-pixel reduction blocks, as the kernel cuts them
    base = 4 * ((Cout + 127) // 128) * max(1, (4 * C) // (128 if (2 * C) % 128 == 0 else 64))
    splits = max(1, min(tiles, 148 // max(base, 1)))
    per = -(-tiles // splits)
    splits = -(-tiles // per)
    part = torch.empty(splits, Cout, 9, 4 * C, dtype=torch.float32, device=x.device)
    if out is None:
        out = torch.empty(Cout, 3, 3, C, dtype=torch.float32, device=x.device)
        accumulate = False
    lib = load()
    with torch.cuda.device(x.device):
        check(lib.scda_conv3x3_s2_wgrad_bf16_nhwc(N, Ho, Wo, C, Cout, x.data_ptr(), dy.data_ptr(), part.data_ptr(),
                                                  splits, stream_ptr(x.device)), "scda_conv3x3_s2_wgrad_bf16_nhwc")
        check(lib.scda_conv_s2_wgrad_gather(Cout, C, part.data_ptr(), splits, out.data_ptr(), 1 if accumulate else 0,
                                            stream_ptr(x.device)), "scda_conv_s2_wgrad_gather")
    return out


def _sink_s2_wgrad(p, x, dy):
    if _direct(p) and _is_krsc(p.grad):
        conv_s2_wgrad(x, dy, out=p.grad, accumulate=not _take_fresh(p))
        return None
    dw = conv_s2_wgrad(x, dy).permute(0, 3, 1, 2)
    if _direct(p):
        p.grad.copy_(dw) if _take_fresh(p) else p.grad.add_(dw)
        return None
    return dw.contiguous()


def _sink_bias_bf16(p, g2d):
    if _direct(p) and p.grad.is_contiguous():
        if _take_fresh(p):
            _clear_fresh(p)
        tc.colsum_into(g2d, p.grad)
        return None
    out = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
    tc.colsum_into(g2d, out)
    return out


# --------------------------------------------------------------------------------------- image discriminator
class _ImageDisFn(torch.autograd.Function):
    """x fp32 [N, 3, H, W] -> logits fp32 [N, (H/8) * (W/8)]: LeakyReLUConv2d x 3 (stride 2) + Conv2d(128, 1, 1)"""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, w4, b4, slope):
        require_cuda(x)
        assert x.dim() == 4 and x.shape[1] == 3 and x.dtype == torch.float32
        N, _, H, W = x.shape
        dev = x.device
        lib = load()
        w1k = w1.detach().permute(0, 2, 3, 1).contiguous()                 # [32,3,3,3] (a view when channels_last)
        y1 = torch.empty(N, H // 2, W // 2, 32, dtype=torch.bfloat16, device=dev)
        with torch.cuda.device(dev):
            check(lib.scda_disc_l1_fwd(N, H, W, x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), x.stride(3),
                                       w1k.data_ptr(), b1.data_ptr(), slope, y1.data_ptr(), 0, stream_ptr(dev)),
                  "scda_disc_l1_fwd")
        y2 = conv_s2(y1, s2_weights(w2), b2.detach(), LEAKY, slope)
        y3 = conv_s2(y2, s2_weights(w3), b3.detach(), LEAKY, slope)
        P, C = y3.numel() // y3.shape[3], y3.shape[3]
        out = torch.empty(N, P // N, dtype=torch.float32, device=dev)
        w4v = w4.detach().reshape(-1).contiguous()
        with torch.cuda.device(dev):
            check(lib.scda_head_dot_fwd(P, C, y3.data_ptr(), 0, w4v.data_ptr(), b4.data_ptr(), out.data_ptr(),
                                        stream_ptr(dev)), "scda_head_dot_fwd")
        ctx.save_for_backward(x, y1, y2, y3)
        ctx.params = (w1, b1, w2, b2, w3, b3, w4, b4)
        ctx.slope = slope
        return out

    @staticmethod
    def backward(ctx, g):
        x, y1, y2, y3 = ctx.saved_tensors
        w1, b1, w2, b2, w3, b3, w4, b4 = ctx.params
        need = ctx.needs_input_grad                     # x, w1, b1, w2, b2, w3, b3, w4, b4, slope
        if FREEZE_PARAMS:
            need = (need[0],) + (False,) * 9
        if FREEZE_INPUT:
            need = (False,) + tuple(need[1:])
        slope, dev, lib = ctx.slope, x.device, load()
        N, _, H, W = x.shape
        P, C = y3.numel() // y3.shape[3], y3.shape[3]
        g = g.contiguous().float()
        grads = [None] * 10
        # head: d3 = gradient w.r.t. the pre-activation of layer 3 (LeakyReLU gradient applied)
        d3 = torch.empty_like(y3)
        w4v = w4.detach().reshape(-1).contiguous()
        dw4 = db4 = None
        if need[7]:
            direct = _direct(w4) and _direct(b4) and w4.grad.is_contiguous()
            if direct:
                if _take_fresh(w4):
                    _clear_fresh(w4)
                if _take_fresh(b4):
                    _clear_fresh(b4)
                dw4, db4 = w4.grad, b4.grad
            else:
                dw4 = torch.zeros(C, dtype=torch.float32, device=dev)
                db4 = torch.zeros(1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.scda_head_dot_bwd(P, C, y3.data_ptr(), 0, w4v.data_ptr(), g.data_ptr(), slope, d3.data_ptr(),
                                        _ptr(dw4), _ptr(db4), stream_ptr(dev)), "scda_head_dot_bwd")
        if need[7] and not (_direct(w4) and _direct(b4) and w4.grad.is_contiguous()):
            grads[7], grads[8] = dw4.view_as(w4), db4.view_as(b4)
        from .gan_ops import run_wgrad_side          # (weight / bias gradients beside the data-gradient chain)
        if need[5]:
            grads[5], grads[6] = run_wgrad_side(
                lambda: (_sink_s2_wgrad(w3, y2, d3), _sink_bias_bf16(b3, d3.view(-1, C))), d3, y2)
        d2 = conv_s2_dgrad(d3, s2_weights(w3), y2.shape[3], mask_src=y2, slope=slope)
        if need[3]:
            grads[3], grads[4] = run_wgrad_side(
                lambda: (_sink_s2_wgrad(w2, y1, d2), _sink_bias_bf16(b2, d2.view(-1, d2.shape[3]))), d2, y1)
        if not (need[0] or need[1]):
            return tuple(grads)
        d1 = conv_s2_dgrad(d2, s2_weights(w2), y1.shape[3], mask_src=y1, slope=slope)
        w1k = w1.detach().permute(0, 2, 3, 1).contiguous()
        dw1 = db1 = dx = ws = None
        wsb = 0
        direct1 = False
        if need[1]:
            direct1 = _direct(w1) and _direct(b1) and _is_krsc(w1.grad)
            acc = 0
            if direct1:
                f1, f2 = _take_fresh(w1), _take_fresh(b1)
                assert f1 == f2
                acc = 0 if f1 else 1
                dw1, db1 = w1.grad, b1.grad
            else:
                dw1 = torch.empty(32, 3, 3, 3, dtype=torch.float32, device=dev)
                db1 = torch.empty(32, dtype=torch.float32, device=dev)
            wsb = lib.scda_disc_l1_workspace_bytes(N, H, W)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        else:
            acc = 0
        if need[0]:
            dx = torch.empty(N, H, W, 3, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.scda_disc_l1_bwd(N, H, W, x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), x.stride(3),
                                       w1k.data_ptr(), d1.data_ptr(), 0, _ptr(dw1), _ptr(db1), _ptr(dx), acc, _ptr(ws),
                                       wsb, stream_ptr(dev)), "scda_disc_l1_bwd")
        if need[1] and not direct1:
            grads[1], grads[2] = dw1.permute(0, 3, 1, 2).contiguous(), db1
        if need[0]:
            grads[0] = dx.permute(0, 3, 1, 2)
        return tuple(grads)


def image_dis_supported(seq, x):
    """the structure _ImageDisFn implements: LeakyReLUConv2d(3, 32) -> (32, 64) -> (64, 128), Conv2d(128, 1, 1)"""
    import torch.nn as nn
    if not (x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and x.shape[1] == 3) or tc.x3():
        return False
    if len(seq) != 4 or not isinstance(seq[3], nn.Conv2d) or seq[3].kernel_size != (1, 1):
        return False
    chans = []
    for blk in list(seq)[:3]:
        m = getattr(blk, "model", None)
        if m is None or len(m) != 2 or not isinstance(m[0], nn.Conv2d) or not isinstance(m[1], nn.LeakyReLU):
            return False
        c = m[0]
        if c.kernel_size != (3, 3) or c.stride != (2, 2) or c.padding != (1, 1) or c.bias is None:
            return False
        chans.append((c.in_channels, c.out_channels))
    return (chans == [(3, 32), (32, 64), (64, 128)] and seq[3].in_channels == 128 and seq[3].out_channels == 1
            and x.shape[2] % 128 == 0 and x.shape[3] % 64 == 0)


def image_dis(seq, x):
    c1, c2, c3, head = seq[0].model[0], seq[1].model[0], seq[2].model[0], seq[3]
    slope = float(seq[0].model[1].negative_slope)
    return _ImageDisFn.apply(x, c1.weight, c1.bias, c2.weight, c2.bias, c3.weight, c3.bias, head.weight, head.bias,
                             slope)


# --------------------------------------------------------------------------------------- feature discriminator
def _bn_fwd(c, bn, slope):
    """c fp32 [N, H, W, C] -> (LeakyReLU(BatchNorm(c)) bf16, mean, rstd); running statistics updated (training)"""
    P, C = c.numel() // c.shape[-1], c.shape[-1]
    y = torch.empty(c.shape, dtype=torch.bfloat16, device=c.device)
    mean = torch.empty(C, dtype=torch.float32, device=c.device)
    rstd = torch.empty(C, dtype=torch.float32, device=c.device)
    mom = 0.1 if bn.momentum is None else float(bn.momentum)
    track = bn.training and bn.track_running_stats
    lib = load()
    wsb = lib.scda_bn_workspace_bytes(P, C)
    ws = torch.empty(wsb, dtype=torch.uint8, device=c.device)
    with torch.cuda.device(c.device):
        check(lib.scda_bn_lrelu_fwd(P, C, c.data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr(), float(bn.eps),
                                    slope, mom, bn.running_mean.data_ptr() if track else None,
                                    bn.running_var.data_ptr() if track else None, mean.data_ptr(), rstd.data_ptr(),
                                    y.data_ptr(), 0, ws.data_ptr(), wsb, stream_ptr(c.device)), "scda_bn_lrelu_fwd")
    if track and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return y, mean, rstd


def _bn_bwd(c, dy, bn, mean, rstd, slope, need_params):
    """-> dc bf16 (gradient w.r.t. the conv output); d gamma / d beta into the sinks"""
    P, C = c.numel() // c.shape[-1], c.shape[-1]
    dc = torch.empty(c.shape, dtype=torch.bfloat16, device=c.device)
    gw, gb = bn.weight, bn.bias
    direct = need_params and _direct(gw) and _direct(gb) and gw.grad.is_contiguous() and gb.grad.is_contiguous()
    if direct:
        f1, f2 = _take_fresh(gw), _take_fresh(gb)
        acc = 0 if (f1 and f2) else 1
        if f1 != f2:
            (gw if f1 else gb).grad.zero_()
        dg, db = gw.grad, gb.grad
    else:
        acc = 0
        dg = torch.empty(C, dtype=torch.float32, device=c.device)
        db = torch.empty(C, dtype=torch.float32, device=c.device)
    lib = load()
    wsb = lib.scda_bn_workspace_bytes(P, C)
    ws = torch.empty(wsb, dtype=torch.uint8, device=c.device)
    with torch.cuda.device(c.device):
        check(lib.scda_bn_lrelu_bwd(P, C, c.data_ptr(), dy.data_ptr(), 1 if dy.dtype == torch.float32 else 0,
                                    gw.data_ptr(), gb.data_ptr(), mean.data_ptr(), rstd.data_ptr(), slope,
                                    dc.data_ptr(), 0, dg.data_ptr(), db.data_ptr(), acc, ws.data_ptr(), wsb,
                                    stream_ptr(c.device)), "scda_bn_lrelu_bwd")
    if not need_params or direct:
        return dc, None, None
    return dc, dg, db


class _PatchDisFn(torch.autograd.Function):
    """x [N, C, 64, 64] (channels-last fp32 or bf16) -> pooled fp32 [N, 4C]: conv s2 - BN - LReLU - conv s2 - BN -
    LReLU - conv s2 - global average pool (ResDis_cluster).  No gradient for x (the cluster features are
    detached, functions/mask.py:234 of the reference)."""

    @staticmethod
    def forward(ctx, x, w1, g1, be1, w2, g2, be2, w3, mods, slope):
        require_cuda(x)
        bn1, bn2 = mods
        # one pass: cast + NCHW -> NHWC (the view below is then contiguous [N, H, W, C])
        xb = x.detach().to(torch.bfloat16, memory_format=torch.channels_last).permute(0, 2, 3, 1)
        c1 = conv_s2(xb, s2_weights(w1), None, 0, out_dtype=torch.float32)
        a1, m1, r1 = _bn_fwd(c1, bn1, slope)
        c2 = conv_s2(a1, s2_weights(w2), None, 0, out_dtype=torch.float32)
        a2, m2, r2 = _bn_fwd(c2, bn2, slope)
        c3 = conv_s2(a2, s2_weights(w3), None, 0, out_dtype=torch.float32)
        N, H3, W3, C3 = c3.shape
        out = torch.empty(N, C3, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(load().scda_avgpool_fwd(N, H3 * W3, C3, c3.data_ptr(), out.data_ptr(), stream_ptr(x.device)),
                  "scda_avgpool_fwd")
        ctx.save_for_backward(xb, c1, a1, m1, r1, c2, a2, m2, r2)
        ctx.params = (w1, w2, w3)
        ctx.mods, ctx.slope, ctx.geom = mods, slope, (N, H3, W3, C3)
        return out

    @staticmethod
    def backward(ctx, g):
        xb, c1, a1, m1, r1, c2, a2, m2, r2 = ctx.saved_tensors
        w1, w2, w3 = ctx.params
        bn1, bn2 = ctx.mods
        need = ctx.needs_input_grad                     # x, w1, g1, be1, w2, g2, be2, w3, mods, slope
        slope = ctx.slope
        N, H3, W3, C3 = ctx.geom
        grads = [None] * 10
        d3 = torch.empty(N, H3, W3, C3, dtype=torch.bfloat16, device=g.device)
        g = g.contiguous().float()
        with torch.cuda.device(g.device):
            check(load().scda_avgpool_bwd(N, H3 * W3, C3, g.data_ptr(), d3.data_ptr(), 0, stream_ptr(g.device)),
                  "scda_avgpool_bwd")
        if need[7]:
            grads[7] = _sink_s2_wgrad(w3, a2, d3)
        da2 = conv_s2_dgrad(d3, s2_weights(w3), a2.shape[3])
        dc2, grads[5], grads[6] = _bn_bwd(c2, da2, bn2, m2, r2, slope, need[5])
        if need[4]:
            grads[4] = _sink_s2_wgrad(w2, a1, dc2)
        da1 = conv_s2_dgrad(dc2, s2_weights(w2), a1.shape[3])
        dc1, grads[2], grads[3] = _bn_bwd(c1, da1, bn1, m1, r1, slope, need[2])
        if need[1]:
            grads[1] = _sink_s2_wgrad(w1, xb, dc1)
        assert not need[0], "the feature discriminator's input (detached cluster features) takes no gradient"
        return tuple(grads)


def patch_dis_supported(seq, x):
    import torch.nn as nn
    if not (x.is_cuda and x.dim() == 4) or tc.x3() or len(seq) != 7:
        return False
    kinds = (nn.Conv2d, nn.BatchNorm2d, nn.LeakyReLU, nn.Conv2d, nn.BatchNorm2d, nn.LeakyReLU, nn.Conv2d)
    if not all(isinstance(m, k) for m, k in zip(seq, kinds)):
        return False
    for c in (seq[0], seq[3], seq[6]):
        if c.kernel_size != (3, 3) or c.stride != (2, 2) or c.padding != (1, 1) or c.bias is not None:
            return False
        if c.in_channels % 32 or c.out_channels % 32:
            return False
    if not (seq[1].training and seq[4].training and seq[1].affine and seq[4].affine):
        return False
    return x.shape[2] % 64 == 0 and x.shape[3] % 64 == 0 and x.shape[1] == seq[0].in_channels \
        and not x.requires_grad


def patch_dis(seq, x):
    slope = float(seq[2].negative_slope)
    return _PatchDisFn.apply(x, seq[0].weight, seq[1].weight, seq[1].bias, seq[3].weight, seq[4].weight, seq[4].bias,
                             seq[6].weight, (seq[1], seq[4]), slope)
