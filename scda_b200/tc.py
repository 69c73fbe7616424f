"""Python entry points of the tcgen05 GEMM / 3x3-convolution kernels (csrc/gemm_tc.cu).

Tensors are bf16 CUDA tensors; activations are NHWC ([N, H, W, C] contiguous), conv
weights [Cout, 3, 3, Cin] ("KRSC").  Outputs are allocated here, the kernels only borrow
pointers (include/scda_b200.h)."""
import torch

from ._lib import check, load, require_cuda, stream_ptr

RELU, OUT_F32, MASK_POS, ACCUMULATE = 1, 2, 4, 8


def gemm_tn(a, b, bias=None, relu=False, out_dtype=torch.bfloat16, mask_src=None, out=None,
            accumulate=False):
    """out[M, N] = a[M, K] @ b[N, K]^T (+ bias) — a, b bf16 row-major (last dim contiguous)."""
    require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1]
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    flags = (RELU if relu else 0) | (OUT_F32 if out.dtype == torch.float32 else 0) \
        | (MASK_POS if mask_src is not None else 0) | (ACCUMULATE if accumulate else 0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    if mask_src is not None:
        assert mask_src.dtype == torch.bfloat16 and mask_src.shape == out.shape \
            and mask_src.stride() == out.stride()
    with torch.cuda.device(a.device):
        check(load().scda_gemm_bf16_tn(M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                                       bias.data_ptr() if bias is not None else None, out.data_ptr(),
                                       out.stride(0), flags,
                                       mask_src.data_ptr() if mask_src is not None else None,
                                       stream_ptr(a.device)), "scda_gemm_bf16_tn")
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
    flags = (RELU if relu else 0) | (OUT_F32 if out_dtype == torch.float32 else 0) \
        | (MASK_POS if mask_src is not None else 0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == Cout and bias.is_contiguous()
    if mask_src is not None:
        assert mask_src.dtype == torch.bfloat16 and mask_src.shape == y.shape and mask_src.is_contiguous()
    with torch.cuda.device(x.device):
        check(load().scda_conv3x3_bf16_nhwc(NB, H, W, Cin, Cout, x.data_ptr(), w_krsc.data_ptr(),
                                            bias.data_ptr() if bias is not None else None, y.data_ptr(),
                                            flags, mask_src.data_ptr() if mask_src is not None else None,
                                            stream_ptr(x.device)), "scda_conv3x3_bf16_nhwc")
    return y


def gemm_nn(a, b, bias=None, relu=False, out_dtype=torch.bfloat16, mask_src=None):
    """out[M, N] = a[M, K] @ b[K, N] — b row-major with N contiguous (e.g. dX = dY @ W)."""
    require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[0]
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[1]
    out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    flags = (RELU if relu else 0) | (OUT_F32 if out_dtype == torch.float32 else 0) \
        | (MASK_POS if mask_src is not None else 0)
    if mask_src is not None:
        assert mask_src.dtype == torch.bfloat16 and mask_src.shape == out.shape and mask_src.is_contiguous()
    with torch.cuda.device(a.device):
        check(load().scda_gemm_bf16_nn(M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                                       bias.data_ptr() if bias is not None else None, out.data_ptr(),
                                       out.stride(0), flags,
                                       mask_src.data_ptr() if mask_src is not None else None,
                                       stream_ptr(a.device)), "scda_gemm_bf16_nn")
    return out


def linear_wgrad(dy, x, out=None):
    """dW[Nout, Kin] fp32 = dy[rows, Nout]^T @ x[rows, Kin]."""
    require_cuda(dy, x)
    assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16
    assert dy.shape[0] == x.shape[0] and dy.stride(1) == 1 and x.stride(1) == 1
    rows, nout = dy.shape
    kin = x.shape[1]
    if out is None:
        out = torch.empty(nout, kin, dtype=torch.float32, device=x.device)
    assert out.dtype == torch.float32 and out.shape == (nout, kin) and out.stride(1) == 1
    with torch.cuda.device(x.device):
        check(load().scda_linear_wgrad_bf16(rows, nout, kin, dy.data_ptr(), dy.stride(0), x.data_ptr(),
                                            x.stride(0), out.data_ptr(), out.stride(0),
                                            stream_ptr(x.device)), "scda_linear_wgrad_bf16")
    return out


def conv3x3_wgrad_nhwc(x, dy, target_ctas=296):
    """dW[Cout, 3, 3, Cin] fp32 from x[N,H,W,Cin], dy[N,H,W,Cout] (both bf16 NHWC)."""
    require_cuda(x, dy)
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16
    assert x.is_contiguous() and dy.is_contiguous() and x.shape[:3] == dy.shape[:3]
    NB, H, W, Cin = x.shape
    Cout = dy.shape[3]
    tiles = NB * H * W // 128
    base = 9 * ((Cout + 127) // 128) * ((Cin + (63 if Cin <= 64 else 127)) // (64 if Cin <= 64 else 128))
    splits = max(1, min(tiles, target_ctas // max(base, 1)))
    per = -(-tiles // splits)
    splits = -(-tiles // per)
    part = torch.empty(splits, Cout, 3, 3, Cin, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(load().scda_conv3x3_wgrad_bf16_nhwc(NB, H, W, Cin, Cout, x.data_ptr(), dy.data_ptr(),
                                                  part.data_ptr(), splits, stream_ptr(x.device)),
              "scda_conv3x3_wgrad_bf16_nhwc")
    return part[0] if splits == 1 else part.sum(0)
