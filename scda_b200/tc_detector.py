"""bf16 NHWC tensor-core execution of the detector's dense layers.

The `VGG` module of models/faster_rcnn/vgg_adver_expansion_cluster.py keeps the reference's
modules and fp32 parameters (names, shapes, state_dict); this runtime executes them:

    features   13 x [conv3x3 + bias + ReLU] (tcgen05 implicit GEMM) with 4 max-pools
               (reference: vgg_adver_expansion_cluster.py:64-65, 101-114)
    rpn_head   conv3x3 + ReLU, then both 1x1 convs as ONE GEMM (models/head.py:20-32)
    rcnn       RoIPool -> fc6 -> fc7 (ReLU + dropout in the GEMM epilogue) -> cls|loc as ONE
               GEMM (vgg_adver_expansion_cluster.py:73-80)

Each stage is one torch.autograd.Function whose backward runs the layer loop by hand: the
data gradient reads the forward weights (no transposed copies), the ReLU gradient rides in
the dgrad epilogue (or in the max-pool backward), the weight gradients are reduced straight
into the optimiser's flat gradient buffer when the parameter opted in
(`p._scda_direct_grad`, set by engine.FlatAdam), otherwise they are returned to autograd.

bf16 weight shadows: `shadow(p)` returns a bf16 copy in the layout the kernels want and
re-derives it whenever `p._version` moved; engine.FlatAdam refreshes the big ones inside
its fused optimiser kernel and stamps `p._scda_shadow_version`.
"""
import os

import torch

from . import tc


def _is_krsc(p):
    """physical layout [O][3][3][I] (channels_last) -> the permuted view is contiguous"""
    return p.dim() == 4 and p.permute(0, 2, 3, 1).is_contiguous()


class _Published(object):
    """a lazily derived tensor set shared by callers on DIFFERENT streams (the source and the target branch of
    the detector forward both use the concatenated head weights): the builder's stream records an event, every
    other stream waits on it before its first use.  Without it the second branch reads the buffer while the
    first is still filling it — a race that stayed hidden as long as one branch was always the slower one."""
    __slots__ = ("value", "event", "stream", "capture")

    @staticmethod
    def _capture_id(stream):
        from ._lib import load
        return int(load().scda_stream_capture_id(stream.cuda_stream))

    def __init__(self, value):
        self.value = value
        self.event = self.stream = None
        self.capture = 0
        if torch.cuda.is_available() and torch.cuda.is_initialized():
            self.stream = torch.cuda.current_stream()
            self.capture = self._capture_id(self.stream)
            self.event = torch.cuda.Event()
            self.event.record(self.stream)

    def get(self):
        if self.event is not None:
            cur = torch.cuda.current_stream()
            # the wait is legal inside the builder's own capture (or outside any); an entry from before a
            # capture is complete (captures begin behind a synchronisation), one from an earlier captured
            # segment is ordered by the engine's segment plan
            if cur != self.stream and self._capture_id(cur) == self.capture:
                cur.wait_event(self.event)
        return self.value


def shadow_of(p):
    """bf16 copy of parameter p: conv weights as [O,3,3,I], everything else as stored.  Kept on
    the parameter (`_scda_shadow`) and re-derived only when `p._version` moved; engine.FlatAdam
    refreshes it inside the optimiser kernel and stamps `_scda_shadow_version` itself."""
    sh = getattr(p, "_scda_shadow", None)
    if sh is not None and getattr(p, "_scda_shadow_version", -1) == p._version:
        return sh
    src = p.detach()
    if p.dim() == 4:
        src = src.permute(0, 2, 3, 1)
    if sh is None or sh.shape != src.shape:
        sh = torch.empty(src.shape, dtype=torch.bfloat16, device=p.device)
        p._scda_shadow = sh
    sh.copy_(src)
    p._scda_shadow_version = p._version
    return sh


def shadow3_of(p, pad_in=None):
    """(fwd, stk) split shadows of parameter p for the 'bf16x3' mode (csrc/x3_ops.cu), cached on the
    parameter and re-derived when it changed (`_version`, or the optimiser's `_scda_epoch`):
    conv [O,I,3,3] -> fwd bf16 [O,3,3,3I] ([hi|hi|lo] per tap), stk bf16 [3O,3,3,I] ([hi;hi;lo]);
    linear [N,K] -> fwd [N,3K], stk [3N,K].  pad_in zero-pads the input channels first (conv1_1)."""
    stamp = (p._version, getattr(p, "_scda_epoch", 0), pad_in)
    ent = getattr(p, "_scda_x3", None)
    if ent is not None and ent[0] == stamp:
        return ent[1].get()
    src = p.detach()
    if p.dim() == 4:
        O, I = p.shape[0], p.shape[1]
        w = src.permute(0, 2, 3, 1)
        if pad_in is not None and pad_in > I:
            wp = torch.zeros(O, 3, 3, pad_in, dtype=torch.float32, device=p.device)
            wp[..., :I].copy_(w)
            w, I = wp, pad_in
        w = w.contiguous()
        fwd, _ = tc.split_weights(w.view(O * 9, I), want_stk=False)
        _, stk = tc.split_weights(w.view(O, 9 * I), want_fwd=False)
        val = (fwd.view(O, 3, 3, 3 * I), stk.view(3 * O, 3, 3, I))
    else:
        val = tc.split_weights(src.contiguous())
    p._scda_x3 = (stamp, _Published(val))
    return val


class TcDetector(object):
    def __init__(self, vgg):
        import torch.nn as nn
        self.convs = []                       # [conv module, pool_after]
        for m in vgg.features.children():
            if isinstance(m, nn.Conv2d):
                assert m.kernel_size == (3, 3) and m.stride == (1, 1) and m.padding == (1, 1)
                self.convs.append([m, False])
            elif isinstance(m, nn.MaxPool2d):
                assert m.kernel_size in (2, (2, 2)) and m.stride in (2, (2, 2)) and self.convs
                self.convs[-1][1] = True
            elif isinstance(m, nn.BatchNorm2d):
                raise NotImplementedError("the tensor-core path covers the non-BN VGG variants")
        self.convs = [tuple(c) for c in self.convs]
        self.vgg = vgg
        self._derived = {}

    # ------------------------------------------------------------------ shadows
    def shadow(self, p):
        return shadow_of(p)

    def _derive(self, key, params, build):
        """small shadows derived from several parameters (padded / concatenated)"""
        stamp = tuple((p._version, getattr(p, "_scda_epoch", 0)) for p in params)
        ent = self._derived.get(key)
        if ent is None or ent[0] != stamp:
            ent = (stamp, _Published(build()))
            self._derived[key] = ent
        return ent[1].get()

    def conv1_weight(self):
        w = self.convs[0][0].weight

        def build():
            out = torch.zeros(w.shape[0], 3, 3, 64, dtype=torch.bfloat16, device=w.device)
            out[..., :w.shape[1]].copy_(w.detach().permute(0, 2, 3, 1))
            return out
        return self._derive("conv1", (w,), build)

    def rpn_cat(self):
        h = self.vgg.rpn_head
        ps = (h.conv_cls.weight, h.conv_cls.bias, h.conv_loc.weight, h.conv_loc.bias)

        def build():
            nc, nl, k = ps[0].shape[0], ps[2].shape[0], ps[0].shape[1]
            n = nc + nl
            w = torch.zeros((n + 7) // 8 * 8, k, dtype=torch.bfloat16, device=ps[0].device)
            w[:nc].copy_(ps[0].detach().view(nc, k))
            w[nc:n].copy_(ps[2].detach().view(nl, k))
            b = torch.cat([ps[1].detach(), ps[3].detach()]).float().contiguous()
            return w, b, nc, nl
        return self._derive("rpn", ps, build)

    def rcnn_cat(self):
        v = self.vgg
        ps = (v.fc_rcnn_cls.weight, v.fc_rcnn_cls.bias, v.fc_rcnn_loc.weight, v.fc_rcnn_loc.bias)

        def build():
            nc, nl, k = ps[0].shape[0], ps[2].shape[0], ps[0].shape[1]
            n = nc + nl
            w = torch.zeros((n + 7) // 8 * 8, k, dtype=torch.bfloat16, device=ps[0].device)
            w[:nc].copy_(ps[0].detach())
            w[nc:n].copy_(ps[2].detach())
            b = torch.cat([ps[1].detach(), ps[3].detach()]).float().contiguous()
            return w, b, nc, nl
        return self._derive("rcnn", ps, build)

    # ------------------------------------------------------------------ fp32-parity mode shadows
    def shadow3(self, p, pad_in=None):
        return shadow3_of(p, pad_in)

    def cat3(self, key):
        """split shadows of the concatenated (cls | loc) head, rows zero-padded to a multiple of 8"""
        if key == "rpn":
            h = self.vgg.rpn_head
            ps = (h.conv_cls.weight, h.conv_cls.bias, h.conv_loc.weight, h.conv_loc.bias)
        else:
            v = self.vgg
            ps = (v.fc_rcnn_cls.weight, v.fc_rcnn_cls.bias, v.fc_rcnn_loc.weight, v.fc_rcnn_loc.bias)

        def build():
            nc, nl, k = ps[0].shape[0], ps[2].shape[0], ps[0].shape[1]
            n = nc + nl
            w = torch.zeros((n + 7) // 8 * 8, k, dtype=torch.float32, device=ps[0].device)
            w[:nc].copy_(ps[0].detach().reshape(nc, k))
            w[nc:n].copy_(ps[2].detach().reshape(nl, k))
            fwd, stk = tc.split_weights(w)
            b = torch.cat([ps[1].detach(), ps[3].detach()]).float().contiguous()
            return fwd, stk, b, nc, nl
        return self._derive("x3" + key, ps, build)

    # ------------------------------------------------------------------ stages
    def features(self, image):
        params = []
        for m, _ in self.convs:
            params += [m.weight, m.bias]
        return (_BackboneFnX3 if tc.x3() else _BackboneFn).apply(image, self, *params)

    def rpn(self, feat):
        h = self.vgg.rpn_head
        return (_RpnHeadFnX3 if tc.x3() else _RpnHeadFn).apply(
            feat, self, h.conv3x3.weight, h.conv3x3.bias, h.conv_cls.weight, h.conv_cls.bias, h.conv_loc.weight,
            h.conv_loc.bias)

    def rcnn(self, feat, rois):
        v = self.vgg
        fc6, fc7 = v.classifier[0], v.classifier[3]
        p_drop = (v.classifier[2].p, v.classifier[5].p) if v.training else (0.0, 0.0)
        pool = v.roipooling
        return (_RcnnHeadFnX3 if tc.x3() else _RcnnHeadFn).apply(feat, rois, self, (pool.pooled_height, pool.pooled_width,
                                                    pool.spatial_scale), p_drop,
                                 fc6.weight, fc6.bias, fc7.weight, fc7.bias,
                                 v.fc_rcnn_cls.weight, v.fc_rcnn_cls.bias,
                                 v.fc_rcnn_loc.weight, v.fc_rcnn_loc.bias)


# ---------------------------------------------------------------------- gradient sinks
def _direct(p):
    return getattr(p, "_scda_direct_grad", False) and p.grad is not None and p.grad.dtype == torch.float32


def _clear_fresh(p):
    """p.grad was found marked fresh: FlatGradBucket.zero() has already zeroed it unless it is one of the large
    gradients it skips (one fill launch per bias gradient saved: ~45 per iteration)"""
    from .utils.distributed_utils import DIRECT_SKIP_NUMEL
    if p.numel() >= DIRECT_SKIP_NUMEL:
        p.grad.zero_()


def _take_fresh(p):
    """True when p.grad is known to hold nothing yet this step (overwrite instead of add)."""
    fresh = getattr(p, "_scda_grad_fresh", False)
    p._scda_grad_fresh = False
    return fresh


def _sink_conv_wgrad(p, x, g, cin_real=None):
    """3x3 weight gradient of parameter p ([O,I,3,3]); returns None when it went straight
    into p.grad, else the gradient tensor for autograd."""
    O, I = p.shape[0], p.shape[1]
    if cin_real is None and _direct(p) and _is_krsc(p.grad):
        tc.conv3x3_wgrad_nhwc(x, g, out=p.grad, accumulate=not _take_fresh(p))
        return None
    dw = tc.conv3x3_wgrad_nhwc(x, g)                       # [O,3,3,Ipad]
    if cin_real is not None:
        dw = dw[..., :cin_real]
    dw = dw.permute(0, 3, 1, 2)
    if _direct(p):
        if _take_fresh(p):
            p.grad.copy_(dw)
        else:
            p.grad.add_(dw)
        return None
    return dw.contiguous()


def _sink_linear_wgrad(p, dy, x):
    if _direct(p) and p.grad.stride(-1) == 1 and p.grad.dim() == 2:
        tc.linear_wgrad(dy, x, out=p.grad, accumulate=not _take_fresh(p))
        return None
    return tc.linear_wgrad(dy, x)


def _sink_small(p, g):
    """small dense gradient g (fp32, shape of p)"""
    if _direct(p):
        if _take_fresh(p):
            p.grad.copy_(g.view_as(p))
        else:
            p.grad.add_(g.view_as(p))
        return None
    return g.view_as(p)


def _sink_bias(p, g2d):
    """bias gradient = column sums of the bf16 gradient matrix g2d [rows, N]"""
    if _direct(p) and p.grad.is_contiguous():
        if _take_fresh(p):
            _clear_fresh(p)
        tc.colsum_into(g2d, p.grad)
        return None
    out = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
    tc.colsum_into(g2d, out)
    return out


# ---------------------------------------------------------------------- backbone
# conv1_1 straight from the fp32 NCHW image (csrc/conv_first.cu); SCDA_FIRST_DIRECT=0 restores the padded form
FIRST_DIRECT = os.environ.get("SCDA_FIRST_DIRECT", "1") != "0"
_PAD_STREAMS = {}


def _pad_stream(device):
    key = torch.device(device).index
    if key not in _PAD_STREAMS:
        _PAD_STREAMS[key] = torch.cuda.Stream(device=device)
    return _PAD_STREAMS[key]


class _BackboneFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, rt, *params):
        assert image.is_cuda and image.dtype == torch.float32 and image.dim() == 4
        image = image.contiguous()
        m0 = rt.convs[0][0]
        direct = FIRST_DIRECT and m0.weight.shape[0] == 64 and m0.weight.shape[1] <= 3
        need_grad = any(ctx.needs_input_grad[2:])
        pad_done = None
        if not direct:
            x = tc.nchw_f32_to_nhwc_bf16(image, 64)
        elif need_grad:
            # conv1_1 reads the fp32 image itself (csrc/conv_first.cu); the zero-padded NHWC copy is only an
            # operand of conv1_1's weight gradient: made on a side stream beside the backbone, joined at its end
            cur = torch.cuda.current_stream()
            side = _pad_stream(image.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                x = tc.nchw_f32_to_nhwc_bf16(image, 64)
                pad_done = torch.cuda.Event()
                pad_done.record(side)
            x.record_stream(cur)
            image.record_stream(side)
        else:
            x = image.new_empty(0)        # (no backward: nothing needs the padded copy)
        saved = [x]
        n = len(rt.convs)
        for i, (m, pool) in enumerate(rt.convs):
            if i == 0 and direct:
                y = tc.conv3x3_first_nchw(image, rt.shadow(m.weight), m.bias.detach(), relu=True)
            else:
                w = rt.conv1_weight() if i == 0 else rt.shadow(m.weight)
                y = tc.conv3x3_nhwc(x, w, m.bias.detach(), relu=True)
            saved.append(y)
            x = tc.maxpool2x2_nhwc(y) if pool else y
            if pool:
                saved.append(x)
        if pad_done is not None:
            torch.cuda.current_stream().wait_event(pad_done)
        ctx.rt = rt
        ctx.cin = image.shape[1]
        ctx.save_for_backward(*saved)
        return x

    @staticmethod
    def backward(ctx, g_feat):
        rt = ctx.rt
        saved = list(ctx.saved_tensors)
        # unpack: input of conv i, output of conv i
        ins, outs = [], []
        k = 0
        x = saved[k]; k += 1
        for (m, pool) in rt.convs:
            ins.append(x)
            y = saved[k]; k += 1
            outs.append(y)
            if pool:
                x = saved[k]; k += 1
            else:
                x = y
        last = len(rt.convs) - 1
        assert not rt.convs[last][1], "the stack ends with a convolution (last pool dropped)"
        g = g_feat.contiguous()
        if g.dtype != torch.bfloat16:
            g = g.to(torch.bfloat16)
        g = torch.where(outs[last] > 0, g, torch.zeros_like(g))       # ReLU of the last conv
        grads = [None] * (2 * len(rt.convs))
        # BODY_WGRAD_STREAM: nothing in the backward reads a layer's weight / bias gradient, and they need only
        # that layer's dY — they go to the side stream, the data-gradient chain (dgrad -> pool -> dgrad ...)
        # stays on the current one; joined at the end
        side = BODY_WGRAD_STREAM if g.is_cuda else None
        cur = torch.cuda.current_stream() if side is not None else None
        hook = BODY_BUCKET_HOOK if g.is_cuda else None
        for i in range(last, -1, -1):
            m, _ = rt.convs[i]
            cout = m.weight.shape[0]
            if hook is not None and i == hook[0] - 1:
                hook[1](torch.cuda.current_stream(), side)
            if side is not None:
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    grads[2 * i + 1] = _sink_bias(m.bias, g.view(-1, cout))
                    grads[2 * i] = _sink_conv_wgrad(m.weight, ins[i], g, cin_real=ctx.cin if i == 0 else None)
                g.record_stream(side)
                if i == 0:
                    break
            else:
                grads[2 * i + 1] = _sink_bias(m.bias, g.view(-1, cout))
                if i == 0:
                    grads[0] = _sink_conv_wgrad(m.weight, ins[0], g, cin_real=ctx.cin)
                    break
                grads[2 * i] = _sink_conv_wgrad(m.weight, ins[i], g)
            w = rt.shadow(m.weight)
            if rt.convs[i - 1][1]:          # the input of conv i is a pooled map
                d_pooled = tc.conv3x3_dgrad_nhwc(g, w)
                g = tc.maxpool2x2_bwd_nhwc(outs[i - 1], d_pooled, relu_mask=True)
            else:
                g = tc.conv3x3_dgrad_nhwc(g, w, mask_src=ins[i])
        if side is not None:
            cur.wait_stream(side)
        return (None, None) + tuple(grads)


# ---------------------------------------------------------------------- RPN head
class _RpnHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, rt, w3, b3, wc, bc, wl, bl):
        NB, H, W, C = feat.shape
        hidden = tc.conv3x3_nhwc(feat, rt.shadow(w3), b3.detach(), relu=True)
        wcat, bcat, nc, nl = rt.rpn_cat()
        out = tc.gemm_tn(hidden.view(-1, hidden.shape[3]), wcat[:nc + nl], bcat, out_dtype=torch.float32)
        out = out.view(NB, H, W, nc + nl)
        cls = out[..., :nc].permute(0, 3, 1, 2).contiguous()
        loc = out[..., nc:].permute(0, 3, 1, 2).contiguous()
        ctx.rt = rt
        ctx.save_for_backward(feat, hidden)
        return cls, loc

    @staticmethod
    def backward(ctx, g_cls, g_loc):
        rt = ctx.rt
        feat, hidden = ctx.saved_tensors
        h = rt.vgg.rpn_head
        w3, b3, wc, bc, wl, bl = (h.conv3x3.weight, h.conv3x3.bias, h.conv_cls.weight, h.conv_cls.bias,
                                  h.conv_loc.weight, h.conv_loc.bias)
        NB, H, W, C = hidden.shape
        wcat, _, nc, nl = rt.rpn_cat()
        npad = wcat.shape[0]
        g = torch.zeros(NB, H, W, npad, dtype=torch.bfloat16, device=feat.device)
        gb = [None, None]
        if g_cls is not None:
            g[..., :nc].copy_(g_cls.permute(0, 2, 3, 1))
            gb[0] = _sink_small(bc, g_cls.sum((0, 2, 3)))
        if g_loc is not None:
            g[..., nc:nc + nl].copy_(g_loc.permute(0, 2, 3, 1))
            gb[1] = _sink_small(bl, g_loc.sum((0, 2, 3)))
        g2 = g.view(-1, npad)
        hid2 = hidden.view(-1, C)
        dwcat = tc.linear_wgrad(g2, hid2)                         # [npad, C] fp32
        gwc = _sink_small(wc, dwcat[:nc])
        gwl = _sink_small(wl, dwcat[nc:nc + nl])
        d_hidden = tc.gemm_nn(g2, wcat, mask_src=hid2).view(NB, H, W, C)
        gb3 = _sink_bias(b3, d_hidden.view(-1, C))
        gw3 = _sink_conv_wgrad(w3, feat, d_hidden)
        d_feat = tc.conv3x3_dgrad_nhwc(d_hidden, rt.shadow(w3))
        return d_feat, None, gw3, gb3, gwc, gb[0], gwl, gb[1]


# ---------------------------------------------------------------------- RCNN head
def _dropout_scale(shape, p, device):
    """keep/scale tensor of nn.Dropout(p) in training: 0 with probability p, else 1/(1-p)."""
    if p <= 0.0:
        return None
    return torch.empty(shape, dtype=torch.bfloat16, device=device).bernoulli_(1.0 - p).mul_(1.0 / (1.0 - p))


# The two keep/scale tensors of the RCNN head do not depend on the RoI features: with MASK_SIDE_STREAMS (set by
# the engine with its stream overlap) they are drawn on a helper stream beside RoIPool instead of between
# RoIPool and fc6 (one helper per issuing stream: the source and the target branch run side by side).
MASK_SIDE_STREAMS = False
_MASK_STREAMS = {}


def _dropout_scales(shapes, ps, device):
    if not MASK_SIDE_STREAMS or not torch.cuda.is_available() or all(p <= 0.0 for p in ps):
        return [_dropout_scale(sh, p, device) for sh, p in zip(shapes, ps)], None
    cur = torch.cuda.current_stream()
    side = _MASK_STREAMS.get(cur.cuda_stream)
    if side is None:
        side = _MASK_STREAMS[cur.cuda_stream] = torch.cuda.Stream(device=device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        out = [_dropout_scale(sh, p, device) for sh, p in zip(shapes, ps)]
    return out, side


def _join_masks(masks, side):
    if side is not None:
        cur = torch.cuda.current_stream()
        cur.wait_stream(side)
        for m in masks:
            if m is not None:
                m.record_stream(cur)


# Side stream for the big fc6 / fc7 weight gradients of the RCNN head's backward (set by the engine; None =
# everything on the current stream).  Those two products are bound by WRITING 478 MB of gradient and nothing in
# the rest of the backward reads them: on the side stream they run beside the data-gradient chain
# (fc7 -> fc6 -> RoIPool backward -> backbone) instead of in front of it.  The engine joins the stream before
# the head bucket's all-reduce / Adam.
WGRAD_STREAM = None
# The same for the backbone's 13 weight / bias gradients (set by the engine with its stream overlap).
BODY_WGRAD_STREAM = None
# (k, fn): fn(main_stream, wgrad_stream_or_None) is called from the backbone's backward once everything of the
# convolutions k .. 12 has been ISSUED (weight / bias gradients and the data gradient that reads conv k's
# weights): the engine reduces and steps that bucket of parameters while the rest of the backward runs.
BODY_BUCKET_HOOK = None


class _RcnnHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, rois, rt, pool, p_drop, w6, b6, w7, b7, wc, bc, wl, bl):
        ph, pw, scale = pool
        NB, H, W, C = feat.shape
        R = rois.shape[0]
        rois = rois.contiguous().float()
        (dm6, dm7), mside = _dropout_scales(((R, w6.shape[0]), (R, w7.shape[0])), p_drop, feat.device)
        x, argmax = tc.roi_pool_nhwc(feat, rois, ph, pw, scale)      # [R, C*ph*pw] bf16, fc6's order
        _join_masks((dm6, dm7), mside)
        h6 = tc.gemm_tn(x, rt.shadow(w6), b6.detach(), relu=True, mul_src=dm6)
        h7 = tc.gemm_tn(h6, rt.shadow(w7), b7.detach(), relu=True, mul_src=dm7)
        wcat, bcat, nc, nl = rt.rcnn_cat()
        out = tc.gemm_tn(h7, wcat[:nc + nl], bcat, out_dtype=torch.float32)
        ctx.rt, ctx.pool, ctx.geom = rt, pool, (NB, H, W, C)
        ctx.has_drop = (dm6 is not None, dm7 is not None)
        keep = [rois, argmax, x, h6, h7] + [t for t in (dm6, dm7) if t is not None]
        ctx.save_for_backward(*keep)
        ctx.set_materialize_grads(False)
        return h7.float(), out[:, :nc].contiguous(), out[:, nc:].contiguous()

    @staticmethod
    def backward(ctx, g_fea, g_cls, g_loc):
        rt = ctx.rt
        saved = list(ctx.saved_tensors)
        rois, argmax, x, h6, h7 = saved[:5]
        rest = saved[5:]
        dm6 = rest.pop(0) if ctx.has_drop[0] else None
        dm7 = rest.pop(0) if ctx.has_drop[1] else None
        v = rt.vgg
        w6, b6, w7, b7 = v.classifier[0].weight, v.classifier[0].bias, v.classifier[3].weight, v.classifier[3].bias
        wc, bc, wl, bl = v.fc_rcnn_cls.weight, v.fc_rcnn_cls.bias, v.fc_rcnn_loc.weight, v.fc_rcnn_loc.bias
        ph, pw, scale = ctx.pool
        NB, H, W, C = ctx.geom
        R = x.shape[0]
        wcat, _, nc, nl = rt.rcnn_cat()
        npad = wcat.shape[0]
        g = torch.zeros(R, npad, dtype=torch.bfloat16, device=x.device)
        gbc = gbl = None
        if g_cls is not None:
            g[:, :nc].copy_(g_cls)
            gbc = _sink_small(bc, g_cls.sum(0))
        if g_loc is not None:
            g[:, nc:nc + nl].copy_(g_loc)
            gbl = _sink_small(bl, g_loc.sum(0))
        dwcat = tc.linear_wgrad(g, h7)
        gwc = _sink_small(wc, dwcat[:nc])
        gwl = _sink_small(wl, dwcat[nc:nc + nl])
        # d(pre-activation of fc7) = (g . Wcat) * dropout scale * [h7 > 0]
        d7 = tc.gemm_nn(g, wcat, mask_src=h7, mul_src=dm7)
        if g_fea is not None:                  # someone differentiates through the fc7 features
            extra = g_fea.to(torch.bfloat16)
            if dm7 is not None:
                extra = extra * dm7
            d7 = d7 + torch.where(h7 > 0, extra, torch.zeros_like(extra))
        gb7 = _sink_bias(b7, d7)
        d6 = tc.gemm_nn(d7, rt.shadow(w7), mask_src=h6, mul_src=dm6)
        gb6 = _sink_bias(b6, d6)
        side = WGRAD_STREAM if (d7.is_cuda and _direct(w6) and _direct(w7)) else None
        if side is not None:
            cur = torch.cuda.current_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                gw7 = _sink_linear_wgrad(w7, d7, h6)
                gw6 = _sink_linear_wgrad(w6, d6, x)
            for t in (d7, d6, h6, x):
                t.record_stream(side)
        else:
            gw7 = _sink_linear_wgrad(w7, d7, h6)
            gw6 = _sink_linear_wgrad(w6, d6, x)
        dx = tc.gemm_nn(d6, rt.shadow(w6))                                    # [R, C*ph*pw] bf16
        d_feat = tc.roi_pool_nhwc_bwd(dx, argmax, rois, (NB, H, W, C), ph, pw).to(torch.bfloat16)
        return d_feat, None, None, None, None, gw6, gb6, gw7, gb7, gwc, gbc, gwl, gbl


# ======================================================================================
# fp32-parity mode ('bf16x3', tc.set_precision): the same three stages with fp32 NHWC activations
# and gradients; every contraction runs on the same tcgen05 kernels with split operands
# (A = [hi|lo|hi], B = [hi|hi|lo]: csrc/x3_ops.cu).  Reference arithmetic: fp32 cuDNN / cuBLAS
# (vgg_adver_expansion_cluster.py:46-60,101-114, models/head.py:13-18).
def _sink_conv_wgrad_x3(p, xs, gs, cin_real=None):
    if cin_real is None and _direct(p) and _is_krsc(p.grad):
        tc.conv3x3_wgrad_x3(xs, gs, out=p.grad, accumulate=not _take_fresh(p))
        return None
    dw = tc.conv3x3_wgrad_x3(xs, gs)                       # [O,3,3,Ipad]
    if cin_real is not None:
        dw = dw[..., :cin_real]
    dw = dw.permute(0, 3, 1, 2)
    if _direct(p):
        if _take_fresh(p):
            p.grad.copy_(dw)
        else:
            p.grad.add_(dw)
        return None
    return dw.contiguous()


def _sink_linear_wgrad_x3(p, gs, xs):
    if _direct(p) and p.grad.stride(-1) == 1 and p.grad.dim() == 2:
        tc.linear_wgrad_x3(gs, xs, out=p.grad, accumulate=not _take_fresh(p))
        return None
    return tc.linear_wgrad_x3(gs, xs)


def _sink_bias_f32(p, g2d):
    if _direct(p) and p.grad.is_contiguous():
        if _take_fresh(p):
            _clear_fresh(p)
        tc.colsum_f32_into(g2d, p.grad)
        return None
    out = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
    tc.colsum_f32_into(g2d, out)
    return out


class _BackboneFnX3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, rt, *params):
        assert image.is_cuda and image.dtype == torch.float32 and image.dim() == 4
        x = tc.nchw_f32_to_nhwc_f32(image.contiguous(), 64)
        saved = []
        for i, (m, pool) in enumerate(rt.convs):
            xs = tc.split3(x)
            wf, _ = rt.shadow3(m.weight, pad_in=64 if i == 0 else None)
            y = tc.conv3x3_nhwc(xs, wf, m.bias.detach(), relu=True, out_dtype=torch.float32)
            saved += [xs, y]
            x = tc.maxpool2x2_nhwc_f32(y) if pool else y
        ctx.rt = rt
        ctx.cin = image.shape[1]
        ctx.save_for_backward(*saved)
        return x

    @staticmethod
    def backward(ctx, g_feat):
        rt = ctx.rt
        saved = list(ctx.saved_tensors)
        ins = saved[0::2]           # split input of conv i
        outs = saved[1::2]          # fp32 output (after ReLU) of conv i
        last = len(rt.convs) - 1
        assert not rt.convs[last][1], "the stack ends with a convolution (last pool dropped)"
        g = g_feat.contiguous().float()
        g = torch.where(outs[last] > 0, g, torch.zeros_like(g))
        grads = [None] * (2 * len(rt.convs))
        for i in range(last, -1, -1):
            m, _ = rt.convs[i]
            cout = m.weight.shape[0]
            grads[2 * i + 1] = _sink_bias_f32(m.bias, g.view(-1, cout))
            gs = tc.split3(g)
            if i == 0:
                grads[0] = _sink_conv_wgrad_x3(m.weight, ins[0], gs, cin_real=ctx.cin)
                break
            grads[2 * i] = _sink_conv_wgrad_x3(m.weight, ins[i], gs)
            _, ws = rt.shadow3(m.weight)
            if rt.convs[i - 1][1]:
                d_pooled = tc.conv3x3_dgrad_nhwc(gs, ws, out_dtype=torch.float32)
                g = tc.maxpool2x2_bwd_nhwc_f32(outs[i - 1], d_pooled, relu_mask=True)
            else:
                g = tc.conv3x3_dgrad_nhwc(gs, ws, mask_src=outs[i - 1], out_dtype=torch.float32)
        return (None, None) + tuple(grads)


class _RpnHeadFnX3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, rt, w3, b3, wc, bc, wl, bl):
        NB, H, W, C = feat.shape
        fs = tc.split3(feat.contiguous().float())
        w3f, _ = rt.shadow3(w3)
        hidden = tc.conv3x3_nhwc(fs, w3f, b3.detach(), relu=True, out_dtype=torch.float32)
        hs = tc.split3(hidden)
        wf, _, bcat, nc, nl = rt.cat3("rpn")
        out = tc.gemm_tn(hs.view(-1, hs.shape[3]), wf[:nc + nl], bcat, out_dtype=torch.float32)
        out = out.view(NB, H, W, nc + nl)
        cls = out[..., :nc].permute(0, 3, 1, 2).contiguous()
        loc = out[..., nc:].permute(0, 3, 1, 2).contiguous()
        ctx.rt = rt
        ctx.save_for_backward(fs, hidden, hs)
        return cls, loc

    @staticmethod
    def backward(ctx, g_cls, g_loc):
        rt = ctx.rt
        fs, hidden, hs = ctx.saved_tensors
        h = rt.vgg.rpn_head
        w3, b3, wc, bc, wl, bl = (h.conv3x3.weight, h.conv3x3.bias, h.conv_cls.weight, h.conv_cls.bias,
                                  h.conv_loc.weight, h.conv_loc.bias)
        NB, H, W, C = hidden.shape
        wf, wstk, _, nc, nl = rt.cat3("rpn")
        npad = wf.shape[0]
        g = torch.zeros(NB, H, W, npad, dtype=torch.float32, device=hidden.device)
        gb = [None, None]
        if g_cls is not None:
            g[..., :nc].copy_(g_cls.permute(0, 2, 3, 1))
            gb[0] = _sink_small(bc, g_cls.sum((0, 2, 3)))
        if g_loc is not None:
            g[..., nc:nc + nl].copy_(g_loc.permute(0, 2, 3, 1))
            gb[1] = _sink_small(bl, g_loc.sum((0, 2, 3)))
        gs = tc.split3(g.view(-1, npad))
        dwcat = tc.linear_wgrad_x3(gs, hs.view(-1, hs.shape[3]))          # [npad, C] fp32
        gwc = _sink_small(wc, dwcat[:nc])
        gwl = _sink_small(wl, dwcat[nc:nc + nl])
        d_hidden = tc.gemm_nn(gs, wstk, mask_src=hidden.view(-1, C), out_dtype=torch.float32).view(NB, H, W, C)
        gb3 = _sink_bias_f32(b3, d_hidden.view(-1, C))
        ds = tc.split3(d_hidden)
        gw3 = _sink_conv_wgrad_x3(w3, fs, ds)
        _, w3s = rt.shadow3(w3)
        d_feat = tc.conv3x3_dgrad_nhwc(ds, w3s, out_dtype=torch.float32)
        return d_feat, None, gw3, gb3, gwc, gb[0], gwl, gb[1]


class _RcnnHeadFnX3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, rois, rt, pool, p_drop, w6, b6, w7, b7, wc, bc, wl, bl):
        ph, pw, scale = pool
        NB, H, W, C = feat.shape
        R = rois.shape[0]
        rois = rois.contiguous().float()
        x, argmax = tc.roi_pool_nhwc_f32(feat.contiguous().float(), rois, ph, pw, scale)    # [R, C*ph*pw] fp32
        dm6 = _dropout_scale((R, w6.shape[0]), p_drop[0], feat.device)
        dm7 = _dropout_scale((R, w7.shape[0]), p_drop[1], feat.device)
        xs = tc.split3(x)
        h6 = tc.gemm_tn(xs, rt.shadow3(w6)[0], b6.detach(), relu=True, mul_src=dm6, out_dtype=torch.float32)
        h6s = tc.split3(h6)
        h7 = tc.gemm_tn(h6s, rt.shadow3(w7)[0], b7.detach(), relu=True, mul_src=dm7, out_dtype=torch.float32)
        h7s = tc.split3(h7)
        wf, _, bcat, nc, nl = rt.cat3("rcnn")
        out = tc.gemm_tn(h7s, wf[:nc + nl], bcat, out_dtype=torch.float32)
        ctx.rt, ctx.pool, ctx.geom = rt, pool, (NB, H, W, C)
        ctx.has_drop = (dm6 is not None, dm7 is not None)
        keep = [rois, argmax, xs, h6, h6s, h7, h7s] + [t for t in (dm6, dm7) if t is not None]
        ctx.save_for_backward(*keep)
        ctx.set_materialize_grads(False)
        return h7, out[:, :nc].contiguous(), out[:, nc:].contiguous()

    @staticmethod
    def backward(ctx, g_fea, g_cls, g_loc):
        rt = ctx.rt
        saved = list(ctx.saved_tensors)
        rois, argmax, xs, h6, h6s, h7, h7s = saved[:7]
        rest = saved[7:]
        dm6 = rest.pop(0) if ctx.has_drop[0] else None
        dm7 = rest.pop(0) if ctx.has_drop[1] else None
        v = rt.vgg
        w6, b6, w7, b7 = v.classifier[0].weight, v.classifier[0].bias, v.classifier[3].weight, v.classifier[3].bias
        wc, bc, wl, bl = v.fc_rcnn_cls.weight, v.fc_rcnn_cls.bias, v.fc_rcnn_loc.weight, v.fc_rcnn_loc.bias
        ph, pw, scale = ctx.pool
        NB, H, W, C = ctx.geom
        R = xs.shape[0]
        wf, wstk, _, nc, nl = rt.cat3("rcnn")
        npad = wf.shape[0]
        g = torch.zeros(R, npad, dtype=torch.float32, device=xs.device)
        gbc = gbl = None
        if g_cls is not None:
            g[:, :nc].copy_(g_cls)
            gbc = _sink_small(bc, g_cls.sum(0))
        if g_loc is not None:
            g[:, nc:nc + nl].copy_(g_loc)
            gbl = _sink_small(bl, g_loc.sum(0))
        gs = tc.split3(g)
        dwcat = tc.linear_wgrad_x3(gs, h7s)
        gwc = _sink_small(wc, dwcat[:nc])
        gwl = _sink_small(wl, dwcat[nc:nc + nl])
        d7 = tc.gemm_nn(gs, wstk, mask_src=h7, mul_src=dm7, out_dtype=torch.float32)
        if g_fea is not None:
            extra = g_fea.float()
            if dm7 is not None:
                extra = extra * dm7.float()
            d7 = d7 + torch.where(h7 > 0, extra, torch.zeros_like(extra))
        gb7 = _sink_bias_f32(b7, d7)
        d7s = tc.split3(d7)
        gw7 = _sink_linear_wgrad_x3(w7, d7s, h6s)
        d6 = tc.gemm_nn(d7s, rt.shadow3(w7)[1], mask_src=h6, mul_src=dm6, out_dtype=torch.float32)
        gb6 = _sink_bias_f32(b6, d6)
        d6s = tc.split3(d6)
        gw6 = _sink_linear_wgrad_x3(w6, d6s, xs)
        dx = tc.gemm_nn(d6s, rt.shadow3(w6)[1], out_dtype=torch.float32)               # [R, C*ph*pw] fp32
        d_feat = tc.roi_pool_nhwc_f32_bwd(dx, argmax, rois, (NB, H, W, C), ph, pw)
        return d_feat, None, None, None, None, gw6, gb6, gw7, gb7, gwc, gbc, gwl, gbl
