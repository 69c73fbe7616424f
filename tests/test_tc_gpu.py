"""tcgen05 GEMM / implicit-GEMM convolution against a plain PyTorch fp32 reference of the
same op on the same bf16-rounded operands.  Tolerance: the kernels accumulate in fp32 and
round the OUTPUT to bf16 (rel 2^-8), so |err| <= 1e-2 * |ref| + 1e-2 * rms(ref)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _close(out, ref, rtol=1e-2):
    import torch
    out, ref = out.float(), ref.float()
    tol = rtol * ref.abs() + rtol * ref.pow(2).mean().sqrt()
    bad = (out - ref).abs() > tol
    assert not bool(bad.any()), "mismatch: %d / %d, max abs err %g (ref rms %g)" % (
        int(bad.sum()), bad.numel(), float((out - ref).abs().max()), float(ref.pow(2).mean().sqrt()))


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (256, 384, 512), (512, 4096, 1024),
                                   (512, 36, 4096), (512, 9, 4096), (300, 4096, 512), (2048, 30, 512),
                                   (130, 70, 72), (512, 4096, 25088)])
def test_gemm_tn(cuda_lib, M, N, K):
    import torch
    from scda_b200 import tc
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ b.float().t() + bias
    out = tc.gemm_tn(a, b, bias)
    _close(out, ref)
    out32 = tc.gemm_tn(a, b, bias, relu=True, out_dtype=torch.float32)
    _close(out32, ref.clamp(min=0), rtol=2e-3)
    # strided A (a row-slice view with a larger leading dimension)
    if K % 16 == 0 and K >= 128:
        big = torch.randn(M, K + 64, device="cuda", generator=g).bfloat16()
        av = big[:, :K]
        _close(tc.gemm_tn(av, b), av.float() @ b.float().t())


def test_gemm_mask_and_accumulate(cuda_lib):
    import torch
    from scda_b200 import tc
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(256, 320, device="cuda", generator=g).bfloat16()
    b = (torch.randn(192, 320, device="cuda", generator=g) / 18).bfloat16()
    mask = torch.randn(256, 192, device="cuda", generator=g).bfloat16()
    ref = a.float() @ b.float().t()
    _close(tc.gemm_tn(a, b, mask_src=mask), torch.where(mask.float() > 0, ref, torch.zeros_like(ref)))
    acc = torch.ones(256, 192, device="cuda")
    tc.gemm_tn(a, b, out=acc, accumulate=True)
    _close(acc, ref + 1, rtol=2e-3)


@pytest.mark.parametrize("NB,H,W,Cin,Cout", [(1, 8, 16, 64, 64), (1, 16, 32, 64, 128), (2, 32, 64, 128, 64),
                                             (1, 32, 64, 512, 512), (1, 64, 128, 256, 512),
                                             (4, 64, 64, 128, 128), (1, 24, 40, 64, 192),
                                             (1, 128, 256, 128, 256)])
def test_conv3x3(cuda_lib, NB, H, W, Cin, Cout):
    import torch
    import torch.nn.functional as F
    from scda_b200 import tc
    g = torch.Generator(device="cuda").manual_seed(H * W + Cin)
    x = torch.randn(NB, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, 3, 3, Cin, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=1)
    ref = ref.permute(0, 2, 3, 1)
    y = tc.conv3x3_nhwc(x, w, bias)
    _close(y, ref)
    y2 = tc.conv3x3_nhwc(x, w, bias, relu=True)
    _close(y2, ref.clamp(min=0))
    assert float(y2.float().min()) >= 0


def test_conv3x3_backbone_full_size(cuda_lib):
    """conv1_2 at the benchmark resolution (64 -> 64 @ 512 x 1024): linearity in the input
    (size-independent property) + a corner / edge spot check against PyTorch."""
    import torch
    import torch.nn.functional as F
    from scda_b200 import tc
    g = torch.Generator(device="cuda").manual_seed(0)
    x1 = torch.randn(1, 512, 1024, 64, device="cuda", generator=g).bfloat16()
    w = (torch.randn(64, 3, 3, 64, device="cuda", generator=g) / 24).bfloat16()
    y1 = tc.conv3x3_nhwc(x1, w, out_dtype=torch.float32)
    y2 = tc.conv3x3_nhwc((x1.float() * 2).bfloat16(), w, out_dtype=torch.float32)   # exact in bf16
    assert float((y2 - 2 * y1).abs().max()) <= 1e-3 * float(y1.abs().max())
    torch.backends.cudnn.allow_tf32 = False
    for hs, ws in ((slice(0, 24), slice(0, 40)), (slice(488, 512), slice(984, 1024))):
        h0 = max(hs.start - 1, 0)
        w0 = max(ws.start - 1, 0)
        patch = x1[:, h0:min(hs.stop + 1, 512), w0:min(ws.stop + 1, 1024)].float().permute(0, 3, 1, 2)
        ref = F.conv2d(patch, w.float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
        ref = ref[:, hs.start - h0:hs.start - h0 + 24, ws.start - w0:ws.start - w0 + 40]
        _close(y1[:, hs, ws], ref, rtol=2e-3)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (512, 25088, 4096), (512, 4096, 36), (300, 512, 72),
                                   (512, 4096, 9)])
def test_gemm_nn(cuda_lib, M, N, K):
    """dX = dY @ W with W read as stored ([out, in], `in` contiguous)."""
    import torch
    from scda_b200 import tc
    if K % 8:
        pytest.skip("leading dimension must be a multiple of 8 elements")
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = (torch.randn(K, N, device="cuda", generator=g) / K ** 0.5).bfloat16()
    _close(tc.gemm_nn(a, b), a.float() @ b.float())


@pytest.mark.parametrize("rows,nout,kin", [(128, 128, 64), (512, 4096, 4096), (512, 4096, 25088),
                                           (300, 256, 192), (512, 40, 4096)])
def test_linear_wgrad(cuda_lib, rows, nout, kin):
    import torch
    from scda_b200 import tc
    if nout % 8:
        pytest.skip("leading dimension must be a multiple of 8 elements")
    g = torch.Generator(device="cuda").manual_seed(rows + nout)
    dy = (torch.randn(rows, nout, device="cuda", generator=g) / rows ** 0.5).bfloat16()
    x = torch.randn(rows, kin, device="cuda", generator=g).bfloat16()
    ref = dy.float().t() @ x.float()
    _close(tc.linear_wgrad(dy, x), ref, rtol=2e-3)


@pytest.mark.parametrize("NB,H,W,Cin,Cout", [(1, 8, 16, 64, 128), (1, 32, 64, 512, 512), (2, 32, 32, 128, 64),
                                             (1, 64, 128, 64, 64), (1, 128, 256, 128, 256), (2, 64, 64, 64, 32)])
def test_conv3x3_wgrad(cuda_lib, NB, H, W, Cin, Cout):
    import torch
    from scda_b200 import tc
    g = torch.Generator(device="cuda").manual_seed(H + W + Cin)
    x = torch.randn(NB, H, W, Cin, device="cuda", generator=g).bfloat16()
    dy = (torch.randn(NB, H, W, Cout, device="cuda", generator=g) / (H * W) ** 0.5).bfloat16()
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, 3, 3),
                                      dy.float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    try:
        for three in (False, True):        # one tap per CTA / one kernel column (three taps) per CTA
            tc.set_wgrad_form(three)
            _close(tc.conv3x3_wgrad_nhwc(x, dy), ref, rtol=3e-3)
            _close(tc.conv3x3_wgrad_nhwc(x, dy, target_ctas=9), ref, rtol=3e-3)       # no split
            # into a gradient buffer: overwrite (the buffer is the slab when nothing is split), then accumulate;
            # both storage forms the sinks pass ([Cout,3,3,Cin] contiguous, channels_last [Cout,Cin,3,3] view)
            o = torch.full((Cout, 3, 3, Cin), 7.0, device="cuda")
            tc.conv3x3_wgrad_nhwc(x, dy, out=o)
            _close(o, ref, rtol=3e-3)
            tc.conv3x3_wgrad_nhwc(x, dy, out=o, accumulate=True)
            _close(o, 2 * ref, rtol=3e-3)
            ocl = torch.full((Cout, 3, 3, Cin), -3.0, device="cuda").permute(0, 3, 1, 2)
            tc.conv3x3_wgrad_nhwc(x, dy, out=ocl)
            _close(ocl.permute(0, 2, 3, 1), ref, rtol=3e-3)
    finally:
        tc.set_wgrad_form(True)             # the default form


def test_conv3x3_dgrad_via_flipped_weights(cuda_lib):
    """The data gradient is the same kernel on flipped / transposed weights, with the ReLU
    gradient of the layer input fused (mask)."""
    import torch
    import torch.nn.functional as F
    from scda_b200 import tc
    g = torch.Generator(device="cuda").manual_seed(3)
    NB, H, W, Cin, Cout = 1, 32, 64, 128, 256
    x = torch.randn(NB, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, 3, 3, Cin, device="cuda", generator=g) / 34).bfloat16()
    dy = torch.randn(NB, H, W, Cout, device="cuda", generator=g).bfloat16()
    wd = w.flip(1, 2).permute(3, 1, 2, 0).contiguous()            # [Cin][3][3][Cout]
    dx = tc.conv3x3_nhwc(dy, wd, mask_src=x)
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.grad.conv2d_input((NB, Cin, H, W), w.float().permute(0, 3, 1, 2),
                                     dy.float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    ref = torch.where(x.float() > 0, ref, torch.zeros_like(ref))
    _close(dx, ref)


def test_linear_wgrad_gemm_form_into_buffer(cuda_lib):
    """the persistent-GEMM form of the big weight gradients (fc6 / fc7): written into, and accumulated onto, a
    caller's fp32 buffer; equal to the one-shot weight-gradient kernel"""
    import torch
    from scda_b200 import tc
    g = torch.Generator(device="cuda").manual_seed(11)
    rows, nout, kin = 512, 512, 8192
    dy = (torch.randn(rows, nout, device="cuda", generator=g) / rows ** 0.5).bfloat16()
    x = torch.randn(rows, kin, device="cuda", generator=g).bfloat16()
    ref = dy.float().t() @ x.float()
    out = torch.full((nout, kin), 7.0, device="cuda")
    assert tc.WGRAD_GEMM
    tc.linear_wgrad(dy, x, out=out)
    _close(out, ref, rtol=2e-3)
    tc.linear_wgrad(dy, x, out=out, accumulate=True)
    _close(out, 2 * ref, rtol=2e-3)
    tc.WGRAD_GEMM = False
    try:
        _close(tc.linear_wgrad(dy, x), out / 2, rtol=1e-3)
    finally:
        tc.WGRAD_GEMM = True


def test_transpose_bf16(cuda_lib):
    import torch
    from scda_b200 import tc
    g = torch.Generator(device="cuda").manual_seed(3)
    for R, C in ((512, 4096), (70, 33), (1, 5)):
        a = torch.randn(R, C + 3, device="cuda", generator=g).bfloat16()[:, :C]       # row stride > C
        assert torch.equal(tc.transpose_bf16(a), a.t().contiguous())
