"""Python entry points of the tcgen05 GEMM / 3x3-convolution kernels (csrc/gemm_tc.cu) and
their bf16 NHWC companions (csrc/nhwc_ops.cu).

Tensors are bf16 CUDA tensors; activations are NHWC ([N, H, W, C] contiguous), conv
weights [Cout, 3, 3, Cin] ("KRSC").  Outputs are allocated here, the kernels only borrow
pointers (include/scda_b200.h)."""
import os

import torch

from ._lib import check, load, require_cuda, stream_ptr

RELU, OUT_F32, MASK_POS, ACCUMULATE, MUL_SRC, MASK_F32 = 1, 2, 4, 8, 16, 64

# Precision mode of the detector / decoder contractions (one process-wide switch):
#   'bf16'    operands rounded to bf16, fp32 accumulation — the throughput mode (default);
#   'bf16x3'  the fp32-parity mode: activations / gradients stay fp32 NHWC, every operand is split
#             into bf16 halves hi + lo and each product is three MMAs (csrc/x3_ops.cu): ~2^-16
#             relative error per product (TF32: 2^-11) at one third of the bf16 rate.  The
#             reference computes these layers in fp32 (cuDNN / cuBLAS of torch 0.4.1).
PRECISION = os.environ.get("SCDA_PRECISION", "bf16")


def set_precision(mode):
    global PRECISION
    if mode not in ("bf16", "bf16x3"):
        raise ValueError("precision mode must be 'bf16' or 'bf16x3', got %r" % (mode,))
    PRECISION = mode
    if mode == "bf16x3":
        # the layers still on cuDNN (discriminators) must not drop to TF32 in the parity mode
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False


def x3():
    return PRECISION == "bf16x3"


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _flags(relu, out_dtype, mask_src, accumulate=False, mul_src=None):
    return (RELU if relu else 0) | (OUT_F32 if out_dtype == torch.float32 else 0) \
        | (MASK_POS if mask_src is not None else 0) | (ACCUMULATE if accumulate else 0) \
        | (MUL_SRC if mul_src is not None else 0) \
        | (MASK_F32 if mask_src is not None and mask_src.dtype == torch.float32 else 0)


def _check_side(t, out, what, f32_ok=False):
    assert t.dtype == torch.bfloat16 or (f32_ok and t.dtype == torch.float32), what
    assert t.shape == out.shape and t.stride() == out.stride(), what


def gemm_tn(a, b, bias=None, relu=False, out_dtype=torch.bfloat16, mask_src=None, out=None,
            accumulate=False, mul_src=None):
    """out[M, N] = a[M, K] @ b[N, K]^T (+ bias) — a, b bf16 row-major (last dim contiguous).
    Epilogue order: bias, ReLU, mask (mask_src > 0), multiply (mul_src)."""
    require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1]
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    flags = _flags(relu, out.dtype, mask_src, accumulate, mul_src)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    if mask_src is not None:
        _check_side(mask_src, out, "mask_src must match the output", f32_ok=True)
    if mul_src is not None:
        _check_side(mul_src, out, "mul_src must match the output")
    with torch.cuda.device(a.device):
        check(load().scda_gemm_bf16_tn(M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                                       _ptr(bias), out.data_ptr(), out.stride(0), flags,
                                       _ptr(mask_src), _ptr(mul_src), stream_ptr(a.device)),
              "scda_gemm_bf16_tn")
    return out


def gemm_nn(a, b, bias=None, relu=False, out_dtype=torch.bfloat16, mask_src=None, mul_src=None, out=None,
            accumulate=False):
    """out[M, N] = a[M, K] @ b[K, N] — b row-major with N contiguous (e.g. dX = dY @ W).
    With `out` (fp32 when accumulate) the result is written / added there."""
    require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[0]
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[1]
    if out is None:
        assert not accumulate
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    else:
        assert out.shape == (M, N) and out.stride(1) == 1
        out_dtype = out.dtype
    flags = _flags(relu, out_dtype, mask_src, accumulate, mul_src)
    if mask_src is not None:
        _check_side(mask_src, out, "mask_src must match the output", f32_ok=True)
    if mul_src is not None:
        _check_side(mul_src, out, "mul_src must match the output")
    with torch.cuda.device(a.device):
        check(load().scda_gemm_bf16_nn(M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                                       _ptr(bias), out.data_ptr(), out.stride(0), flags,
                                       _ptr(mask_src), _ptr(mul_src), stream_ptr(a.device)),
              "scda_gemm_bf16_nn")
    return out


def conv3x3_nhwc(x, w_krsc, bias=None, relu=False, out_dtype=torch.bfloat16, mask_src=None):
    """y[N,H,W,Cout] = conv3x3(x[N,H,W,Cin], w[Cout,3,3,Cin]), stride 1, zero padding 1."""
    require_cuda(x, w_krsc)
    assert x.dtype == torch.bfloat16 and w_krsc.dtype == torch.bfloat16
    assert x.is_contiguous() and w_krsc.is_contiguous() and x.dim() == 4 and w_krsc.dim() == 4
    NB, H, W, Cin = x.shape
    Cout = w_krsc.shape[0]
    assert tuple(w_krsc.shape[1:]) == (3, 3, Cin)
    y = torch.empty(NB, H, W, Cout, dtype=out_dtype, device=x.device)
    flags = _flags(relu, out_dtype, mask_src)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == Cout and bias.is_contiguous()
    if mask_src is not None:
        assert mask_src.dtype in (torch.bfloat16, torch.float32) and mask_src.shape == y.shape \
            and mask_src.is_contiguous()
    with torch.cuda.device(x.device):
        check(load().scda_conv3x3_bf16_nhwc(NB, H, W, Cin, Cout, x.data_ptr(), w_krsc.data_ptr(),
                                            _ptr(bias), y.data_ptr(), flags, _ptr(mask_src),
                                            stream_ptr(x.device)), "scda_conv3x3_bf16_nhwc")
    return y


def conv3x3_first_nchw(image, w_krsc, bias, relu=True):
    """first backbone convolution straight from the fp32 NCHW image (scda_conv3x3_first_nchw): image
    [NB,Cin<=3,H,W] fp32, w_krsc [64,3,3,Cin] bf16, bias [64] fp32 -> [NB,H,W,64] bf16 (+ReLU)"""
    require_cuda(image, w_krsc, bias)
    assert image.dtype == torch.float32 and image.is_contiguous() and image.dim() == 4
    NB, Cin, H, W = image.shape
    Cout = w_krsc.shape[0]
    assert w_krsc.dtype == torch.bfloat16 and w_krsc.is_contiguous() and tuple(w_krsc.shape) == (Cout, 3, 3, Cin)
    assert bias.dtype == torch.float32 and bias.is_contiguous()
    y = torch.empty(NB, H, W, Cout, dtype=torch.bfloat16, device=image.device)
    with torch.cuda.device(image.device):
        check(load().scda_conv3x3_first_nchw(NB, H, W, Cin, Cout, image.data_ptr(), w_krsc.data_ptr(),
                                             bias.data_ptr(), y.data_ptr(), int(bool(relu)),
                                             stream_ptr(image.device)), "scda_conv3x3_first_nchw")
    return y


def set_conv_plan(halo=-1, block_n=-1, sub_tiles=-1):
    """tuning / test hook (scda_conv3x3_set_plan): halo 1|0, block_n 0|64|128|256 (256: pairs only), sub_tiles 0|1|2"""
    if load().scda_conv3x3_set_plan(int(halo), int(block_n), int(sub_tiles)) != 1:
        raise ValueError("scda_conv3x3_set_plan: bad arguments")


def set_conv_pair(mode=-1):
    """tuning / test hook (scda_conv3x3_set_pair): CTA-pair form of the halo kernel, -1 plan | 0 never | 1 always"""
    if load().scda_conv3x3_set_pair(int(mode)) != 1:
        raise ValueError("scda_conv3x3_set_pair: bad argument")


def conv3x3_dgrad_nhwc(dy, w_krsc, mask_src=None, out_dtype=torch.bfloat16):
    """dx[N,H,W,Cin] from dy[N,H,W,Cout] and the FORWARD weights w[Cout,3,3,Cin]; with
    mask_src (= the layer input, a ReLU output) the ReLU gradient is applied as well."""
    require_cuda(dy, w_krsc)
    assert dy.dtype == torch.bfloat16 and w_krsc.dtype == torch.bfloat16
    assert dy.is_contiguous() and w_krsc.is_contiguous() and dy.dim() == 4 and w_krsc.dim() == 4
    NB, H, W, Cout = dy.shape
    Cin = w_krsc.shape[3]
    assert tuple(w_krsc.shape[:3]) == (Cout, 3, 3)
    dx = torch.empty(NB, H, W, Cin, dtype=out_dtype, device=dy.device)
    flags = _flags(False, out_dtype, mask_src)
    if mask_src is not None:
        assert mask_src.dtype in (torch.bfloat16, torch.float32) and mask_src.shape == dx.shape \
            and mask_src.is_contiguous()
    with torch.cuda.device(dy.device):
        check(load().scda_conv3x3_dgrad_bf16_nhwc(NB, H, W, Cin, Cout, dy.data_ptr(), w_krsc.data_ptr(),
                                                  dx.data_ptr(), flags, _ptr(mask_src),
                                                  stream_ptr(dy.device)), "scda_conv3x3_dgrad_bf16_nhwc")
    return dx


WGRAD_GEMM = os.environ.get("SCDA_WGRAD_GEMM", "1") != "0"


def transpose_bf16(a):
    """a bf16 [R, C] (row stride >= C) -> contiguous [C, R]"""
    require_cuda(a)
    assert a.dtype == torch.bfloat16 and a.dim() == 2 and a.stride(1) == 1
    R, Cn = a.shape
    out = torch.empty(Cn, R, dtype=torch.bfloat16, device=a.device)
    with torch.cuda.device(a.device):
        check(load().scda_transpose_bf16(R, Cn, a.data_ptr(), a.stride(0), out.data_ptr(), R, stream_ptr(a.device)),
              "scda_transpose_bf16")
    return out


def linear_wgrad(dy, x, out=None, accumulate=False):
    """dW[Nout, Kin] fp32 (+)= dy[rows, Nout]^T @ x[rows, Kin]."""
    require_cuda(dy, x)
    assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16
    assert dy.shape[0] == x.shape[0] and dy.stride(1) == 1 and x.stride(1) == 1
    rows, nout = dy.shape
    kin = x.shape[1]
    if (WGRAD_GEMM and rows <= 2048 and rows % 64 == 0 and nout % 128 == 0 and kin % 256 == 0
            and nout * kin >= (1 << 22) and x.stride(0) % 8 == 0
            and (out is None or (out.dtype == torch.float32 and out.stride(1) == 1 and out.stride(0) % 8 == 0))):
        # short reduction, big output (fc6: 512 rows -> 4096 x 25088 = 411 MB of fp32): the product is bound by
        # WRITING the gradient.  The one-shot tc_wgrad_kernel (6272 CTAs of four k-blocks, no overlap of a tile's
        # store with the next tile's MMAs) took 0.31 ms alone and 1.2 ms inside the iteration; the persistent
        # GEMM kernel (256-wide tiles, double-buffered TMEM accumulators) does it on dY transposed
        # (4 MB: one small copy) as out = dY^T[Nout, rows] . X[rows, Kin].
        return gemm_nn(transpose_bf16(dy), x, out_dtype=torch.float32, out=out, accumulate=accumulate)
    if out is None:
        assert not accumulate
        out = torch.empty(nout, kin, dtype=torch.float32, device=x.device)
    assert out.dtype == torch.float32 and out.shape == (nout, kin) and out.stride(1) == 1
    with torch.cuda.device(x.device):
        check(load().scda_linear_wgrad_bf16(rows, nout, kin, dy.data_ptr(), dy.stride(0), x.data_ptr(),
                                            x.stride(0), out.data_ptr(), out.stride(0),
                                            1 if accumulate else 0, stream_ptr(x.device)),
              "scda_linear_wgrad_bf16")
    return out


# three taps (one kernel column) per CTA (csrc/gemm_tc.cu: tc_wgrad3_kernel): the default.  Per layer 5-30 % faster than
# the one-tap form (conv1_2 116 -> 81 us, conv3_2 / conv4_2 58 -> 50 us) and 7.00 -> 6.84 ms per iteration
# (gpurun r2_u); SCDA_WGRAD3=0 selects the one-tap form.
WGRAD3 = os.environ.get("SCDA_WGRAD3", "1") == "1"


SHORT_K_ONE_TAP = os.environ.get("SCDA_WGRAD_SHORT_K", "1") == "1"


def _wgrad_splits(NB, H, W, Cin, Cout, target_ctas):
    """how many ranges the pixel reduction is cut into: enough CTAs for `target_ctas`, every
    range non-empty.  CTAs per range = (taps or kernel columns) x Cout tiles x Cin tiles."""
    tiles = NB * H * W // 128
    co_t, ci_t = (Cout + 127) // 128, (Cin + (63 if Cin <= 64 else 127)) // (64 if Cin <= 64 else 128)
    if WGRAD3 and SHORT_K_ONE_TAP and tiles <= 16 and 9 * co_t * ci_t >= 96:
        # short reduction, big output (conv5_x / RPN: 16 pixel tiles, 512 x 4608 outputs): one tap per CTA fills
        # the GPU without any split, so no slab is written and re-read (the C side picks that form when
        # splits == 1 on such a shape; cuDNN's weight gradient was 2.2x faster here, profiles/r2_final_convbench.jsonl)
        return 1
    base = (3 if WGRAD3 else 9) * co_t * ci_t
    splits = max(1, min(tiles, target_ctas // max(base, 1)))
    per = -(-tiles // splits)
    return -(-tiles // per)


def set_wgrad_form(three_taps):
    """tuning / test hook (scda_conv3x3_wgrad_set_form): one tap (False) or one kernel column (True) per CTA"""
    global WGRAD3
    if load().scda_conv3x3_wgrad_set_form(1 if three_taps else 0) != 1:
        raise ValueError("scda_conv3x3_wgrad_set_form: bad arguments")
    WGRAD3 = bool(three_taps)


def conv3x3_wgrad_nhwc(x, dy, target_ctas=None, out=None, accumulate=False):
    """dW[Cout, 3, 3, Cin] fp32 from x[N,H,W,Cin], dy[N,H,W,Cout] (both bf16 NHWC).  With
    `out` (a contiguous fp32 [Cout,3,3,Cin] buffer) the split-K slabs are reduced straight
    into it (overwriting, or adding when accumulate)."""
    require_cuda(x, dy)
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16
    assert x.is_contiguous() and dy.is_contiguous() and x.shape[:3] == dy.shape[:3]
    NB, H, W, Cin = x.shape
    Cout = dy.shape[3]
    if target_ctas is None:
        target_ctas = 148 if WGRAD3 else 296          # one wave of one-CTA-per-SM blocks / two waves
    splits = _wgrad_splits(NB, H, W, Cin, Cout, target_ctas)
    if out is not None:
        assert out.dtype == torch.float32 and out.numel() == Cout * 9 * Cin
    # one slab that overwrites: the gradient buffer itself is the slab (its storage is [Cout][3][3][Cin]: a
    # contiguous [Cout,3,3,Cin] tensor or a channels_last [Cout,Cin,3,3] view of the flat gradient)
    dense = out is not None and ((tuple(out.shape) == (Cout, 3, 3, Cin) and out.is_contiguous())
                                 or (tuple(out.shape) == (Cout, Cin, 3, 3)
                                     and out.is_contiguous(memory_format=torch.channels_last)))
    direct = splits == 1 and dense and not accumulate
    part = None if direct else torch.empty(splits, Cout, 3, 3, Cin, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(load().scda_conv3x3_wgrad_bf16_nhwc(NB, H, W, Cin, Cout, x.data_ptr(), dy.data_ptr(),
                                                  out.data_ptr() if direct else part.data_ptr(), splits,
                                                  stream_ptr(x.device)),
              "scda_conv3x3_wgrad_bf16_nhwc")
    if out is None:
        return part[0] if splits == 1 else part.sum(0)
    if not direct:
        reduce_slabs(part, out, accumulate)
    return out


def reduce_slabs(part, out, accumulate=False):
    """out (+)= part.sum(0); `out` is any fp32 tensor whose storage order matches a slab."""
    n = part[0].numel()
    with torch.cuda.device(part.device):
        check(load().scda_reduce_slabs_f32(part.data_ptr(), n, part.shape[0], out.data_ptr(), n,
                                           1 if accumulate else 0, stream_ptr(part.device)),
              "scda_reduce_slabs_f32")
    return out


def maxpool2x2_nhwc(x):
    require_cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    NB, H, W, C = x.shape
    y = torch.empty(NB, H // 2, W // 2, C, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(load().scda_maxpool2x2_nhwc_bf16(NB, H, W, C, x.data_ptr(), y.data_ptr(),
                                               stream_ptr(x.device)), "scda_maxpool2x2_nhwc_bf16")
    return y


def maxpool2x2_bwd_nhwc(x, dy, relu_mask=True):
    """gradient of maxpool2x2_nhwc at x; with relu_mask also through the ReLU that produced x."""
    require_cuda(x, dy)
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and x.is_contiguous() and dy.is_contiguous()
    NB, H, W, C = x.shape
    assert tuple(dy.shape) == (NB, H // 2, W // 2, C)
    dx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(load().scda_maxpool2x2_bwd_nhwc_bf16(NB, H, W, C, x.data_ptr(), dy.data_ptr(), dx.data_ptr(),
                                                   1 if relu_mask else 0, stream_ptr(x.device)),
              "scda_maxpool2x2_bwd_nhwc_bf16")
    return dx


def nchw_f32_to_nhwc_bf16(x, c_pad=None):
    require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    NB, C, H, W = x.shape
    c_pad = C if c_pad is None else c_pad
    y = torch.empty(NB, H, W, c_pad, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(load().scda_nchw_f32_to_nhwc_bf16(NB, C, H, W, c_pad, x.data_ptr(), y.data_ptr(),
                                                stream_ptr(x.device)), "scda_nchw_f32_to_nhwc_bf16")
    return y


def nhwc_bf16_to_nchw_f32(x):
    require_cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    NB, H, W, C = x.shape
    y = torch.empty(NB, C, H, W, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(load().scda_nhwc_bf16_to_nchw_f32(NB, C, H, W, x.data_ptr(), y.data_ptr(),
                                                stream_ptr(x.device)), "scda_nhwc_bf16_to_nchw_f32")
    return y


def colsum_into(x2d, out):
    """out[N] (fp32) += column sums of x2d[M, N] (bf16, row stride any even number)."""
    require_cuda(x2d, out)
    assert x2d.dtype == torch.bfloat16 and x2d.dim() == 2 and x2d.stride(1) == 1
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == x2d.shape[1]
    with torch.cuda.device(x2d.device):
        check(load().scda_colsum_bf16(x2d.shape[0], x2d.shape[1], x2d.data_ptr(), x2d.stride(0),
                                      out.data_ptr(), stream_ptr(x2d.device)), "scda_colsum_bf16")
    return out


def roi_pool_nhwc(feat, rois, pooled_height, pooled_width, spatial_scale):
    """feat [NB,H,W,C] bf16, rois [R,5] fp32 -> (out [R, C*PH*PW] bf16 in channel-major order,
    argmax uint16 view as int16 tensor of the same shape)."""
    require_cuda(feat, rois)
    assert feat.dtype == torch.bfloat16 and feat.is_contiguous() and feat.dim() == 4
    assert rois.dtype == torch.float32 and rois.is_contiguous() and rois.dim() == 2 and rois.shape[1] == 5
    NB, H, W, C = feat.shape
    R = rois.shape[0]
    n = C * pooled_height * pooled_width
    out = torch.empty(R, n, dtype=torch.bfloat16, device=feat.device)
    argmax = torch.empty(R, n, dtype=torch.int16, device=feat.device)      # bits of a uint16
    with torch.cuda.device(feat.device):
        check(load().scda_roi_pool_nhwc_bf16_fwd(feat.data_ptr(), spatial_scale, R, NB, H, W, C, pooled_height,
                                                 pooled_width, rois.data_ptr(), out.data_ptr(),
                                                 argmax.data_ptr(), stream_ptr(feat.device)),
              "scda_roi_pool_nhwc_bf16_fwd")
    return out, argmax


def roi_pool_nhwc_bwd(dout, argmax, rois, geom, pooled_height, pooled_width):
    """dout [R, C*PH*PW] bf16 -> dfeat [NB,H,W,C] fp32"""
    require_cuda(dout, argmax, rois)
    NB, H, W, C = geom
    assert dout.dtype == torch.bfloat16 and dout.is_contiguous() and argmax.is_contiguous()
    dfeat = torch.empty(NB, H, W, C, dtype=torch.float32, device=dout.device)
    with torch.cuda.device(dout.device):
        check(load().scda_roi_pool_nhwc_bf16_bwd(dout.data_ptr(), argmax.data_ptr(), rois.data_ptr(),
                                                 rois.shape[0], NB, H, W, C, pooled_height, pooled_width,
                                                 dfeat.data_ptr(), stream_ptr(dout.device)),
              "scda_roi_pool_nhwc_bf16_bwd")
    return dfeat


# ---------------------------------------------------------------------------------------
# fp32-parity mode (PRECISION == 'bf16x3'): operand splits and fp32 NHWC companions (csrc/x3_ops.cu)
def split3(x):
    """x fp32 [..., C] (last dim contiguous, uniform row stride) -> bf16 [..., 3C] = [hi | lo | hi]"""
    require_cuda(x)
    assert x.dtype == torch.float32 and x.stride(-1) == 1
    C = x.shape[-1]
    if x.dim() == 2:
        rows, ld = x.shape[0], x.stride(0)
    else:
        assert x.is_contiguous()
        rows, ld = x.numel() // C, C
    y = torch.empty(x.shape[:-1] + (3 * C,), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(load().scda_split3_f32_bf16(rows, C, x.data_ptr(), ld, y.data_ptr(), stream_ptr(x.device)),
              "scda_split3_f32_bf16")
    return y


def split_weights(w2d, want_fwd=True, want_stk=True):
    """w fp32 [rows, K] contiguous -> (fwd bf16 [rows, 3K] = [hi|hi|lo], stk bf16 [3 rows, K] = [hi;hi;lo])"""
    require_cuda(w2d)
    assert w2d.dtype == torch.float32 and w2d.dim() == 2 and w2d.is_contiguous()
    rows, K = w2d.shape
    fwd = torch.empty(rows, 3 * K, dtype=torch.bfloat16, device=w2d.device) if want_fwd else None
    stk = torch.empty(3 * rows, K, dtype=torch.bfloat16, device=w2d.device) if want_stk else None
    with torch.cuda.device(w2d.device):
        check(load().scda_split_weights_f32_bf16(rows, K, w2d.data_ptr(), K, _ptr(fwd), _ptr(stk),
                                                 stream_ptr(w2d.device)), "scda_split_weights_f32_bf16")
    return fwd, stk


def conv3x3_wgrad_x3(xs, gs, out=None, accumulate=False, target_ctas=None):
    """dW[Cout,3,3,Cin] fp32 from the SPLIT operands xs bf16 [N,H,W,3Cin] = [hi|lo|hi] and gs bf16
    [N,H,W,3Cout]: x_hi g_hi + x_lo g_hi + x_hi g_lo, three launches of the strided weight-gradient
    kernel into separate split-K slabs, one slab reduction."""
    require_cuda(xs, gs)
    assert xs.dtype == torch.bfloat16 and gs.dtype == torch.bfloat16 and xs.is_contiguous() and gs.is_contiguous()
    NB, H, W, C3 = xs.shape
    Cin, Cout = C3 // 3, gs.shape[3] // 3
    if target_ctas is None:
        target_ctas = 148 if WGRAD3 else 296
    splits = _wgrad_splits(NB, H, W, Cin, Cout, target_ctas)
    part = torch.empty(3 * splits, Cout, 3, 3, Cin, dtype=torch.float32, device=xs.device)
    lib = load()
    pairs = ((0, 0), (1, 0), (0, 1))              # (x block, g block): hi.hi, lo.hi, hi.lo
    with torch.cuda.device(xs.device):
        for k, (bx, bg) in enumerate(pairs):
            check(lib.scda_conv3x3_wgrad_bf16_nhwc_ld(
                NB, H, W, Cin, Cout, xs.data_ptr() + 2 * bx * Cin, 3 * Cin, gs.data_ptr() + 2 * bg * Cout, 3 * Cout,
                part[k * splits].data_ptr(), splits, stream_ptr(xs.device)), "scda_conv3x3_wgrad_bf16_nhwc_ld")
    if out is None:
        return part.sum(0)
    assert out.dtype == torch.float32 and out.numel() == part[0].numel()
    reduce_slabs(part, out, accumulate)
    return out


def linear_wgrad_x3(gs, xs, out=None, accumulate=False):
    """dW[Nout, Kin] fp32 (+)= dY^T X from the split operands gs bf16 [rows, 3Nout], xs bf16 [rows, 3Kin]"""
    nout, kin = gs.shape[1] // 3, xs.shape[1] // 3
    if out is None:
        assert not accumulate
        out = torch.empty(nout, kin, dtype=torch.float32, device=xs.device)
    for k, (bg, bx) in enumerate(((0, 0), (0, 1), (1, 0))):
        linear_wgrad(gs[:, bg * nout:(bg + 1) * nout], xs[:, bx * kin:(bx + 1) * kin], out=out,
                     accumulate=accumulate or k > 0)
    return out


def maxpool2x2_nhwc_f32(x):
    require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    NB, H, W, C = x.shape
    y = torch.empty(NB, H // 2, W // 2, C, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(load().scda_maxpool2x2_nhwc_f32(NB, H, W, C, x.data_ptr(), y.data_ptr(), stream_ptr(x.device)),
              "scda_maxpool2x2_nhwc_f32")
    return y


def maxpool2x2_bwd_nhwc_f32(x, dy, relu_mask=True):
    require_cuda(x, dy)
    assert x.dtype == torch.float32 and dy.dtype == torch.float32 and x.is_contiguous() and dy.is_contiguous()
    NB, H, W, C = x.shape
    assert tuple(dy.shape) == (NB, H // 2, W // 2, C)
    dx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(load().scda_maxpool2x2_bwd_nhwc_f32(NB, H, W, C, x.data_ptr(), dy.data_ptr(), dx.data_ptr(),
                                                  1 if relu_mask else 0, stream_ptr(x.device)),
              "scda_maxpool2x2_bwd_nhwc_f32")
    return dx


def nchw_f32_to_nhwc_f32(x, c_pad=None):
    require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    NB, C, H, W = x.shape
    c_pad = C if c_pad is None else c_pad
    y = torch.empty(NB, H, W, c_pad, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(load().scda_nchw_f32_to_nhwc_f32(NB, C, H, W, c_pad, x.data_ptr(), y.data_ptr(),
                                               stream_ptr(x.device)), "scda_nchw_f32_to_nhwc_f32")
    return y


def colsum_f32_into(x2d, out):
    """out[N] (fp32) += column sums of x2d[M, N] (fp32, any row stride)."""
    require_cuda(x2d, out)
    assert x2d.dtype == torch.float32 and x2d.dim() == 2 and x2d.stride(1) == 1
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == x2d.shape[1]
    with torch.cuda.device(x2d.device):
        check(load().scda_colsum_f32_ld(x2d.shape[0], x2d.shape[1], x2d.data_ptr(), x2d.stride(0),
                                        out.data_ptr(), stream_ptr(x2d.device)), "scda_colsum_f32_ld")
    return out


def roi_pool_nhwc_f32(feat, rois, pooled_height, pooled_width, spatial_scale):
    """feat [NB,H,W,C] fp32, rois [R,5] fp32 -> (out [R, C*PH*PW] fp32 channel-major, argmax int16 bits)"""
    require_cuda(feat, rois)
    assert feat.dtype == torch.float32 and feat.is_contiguous() and feat.dim() == 4
    assert rois.dtype == torch.float32 and rois.is_contiguous() and rois.dim() == 2 and rois.shape[1] == 5
    NB, H, W, C = feat.shape
    R = rois.shape[0]
    n = C * pooled_height * pooled_width
    out = torch.empty(R, n, dtype=torch.float32, device=feat.device)
    argmax = torch.empty(R, n, dtype=torch.int16, device=feat.device)
    with torch.cuda.device(feat.device):
        check(load().scda_roi_pool_nhwc_f32_fwd(feat.data_ptr(), spatial_scale, R, NB, H, W, C, pooled_height,
                                                pooled_width, rois.data_ptr(), out.data_ptr(),
                                                argmax.data_ptr(), stream_ptr(feat.device)),
              "scda_roi_pool_nhwc_f32_fwd")
    return out, argmax


def roi_pool_nhwc_f32_bwd(dout, argmax, rois, geom, pooled_height, pooled_width):
    require_cuda(dout, argmax, rois)
    NB, H, W, C = geom
    assert dout.dtype == torch.float32 and dout.is_contiguous() and argmax.is_contiguous()
    dfeat = torch.empty(NB, H, W, C, dtype=torch.float32, device=dout.device)
    with torch.cuda.device(dout.device):
        check(load().scda_roi_pool_nhwc_f32_bwd(dout.data_ptr(), argmax.data_ptr(), rois.data_ptr(),
                                                rois.shape[0], NB, H, W, C, pooled_height, pooled_width,
                                                dfeat.data_ptr(), stream_ptr(dout.device)),
              "scda_roi_pool_nhwc_f32_bwd")
    return dfeat
