#!/usr/bin/env python
"""Per-layer timing of the tcgen05 kernels at the VGG16 / head shapes of the 512x1024
benchmark image: forward, data gradient, weight gradient -> TFLOP/s and fraction of the
measured dense bf16 peak (MEASURED_PEAKS.json).  CUDA events on the launch stream, L2
flushed between timed launches.  Writes one JSON line per (layer, pass)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scda_b200 import tc  # noqa: E402

LAYERS = [  # name, H, W, Cin, Cout
    ("conv1_1p", 512, 1024, 64, 64), ("conv1_2", 512, 1024, 64, 64),
    ("conv2_1", 256, 512, 64, 128), ("conv2_2", 256, 512, 128, 128),
    ("conv3_1", 128, 256, 128, 256), ("conv3_2", 128, 256, 256, 256),
    ("conv4_1", 64, 128, 256, 512), ("conv4_2", 64, 128, 512, 512),
    ("conv5_x", 32, 64, 512, 512),
]
GEMMS = [  # name, M, N, K
    ("fc6", 512, 4096, 25088), ("fc7", 512, 4096, 4096), ("rpn_1x1", 2048, 90, 512),
]


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["bf16_tflops"] if os.path.exists(p) else 1590.0


def timeit(fn, flush, iters=6, warm=2):
    ts = []
    for i in range(warm + iters):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    """`lib_us`: the same product through the library the reference would use today — cuDNN (bf16, channels_last,
    benchmark mode: its best algorithm) for the convolutions, cuBLAS (torch.matmul, bf16) for the fc layers —
    timed the same way.  The library calls compute the bare product (no bias / ReLU / mask epilogue, no split-K
    slab reduction beyond their own)."""
    import torch.nn.functional as F
    dev = torch.device("cuda")
    pk = peak()
    torch.backends.cudnn.benchmark = True
    flush = torch.zeros(64 * 1024 * 1024, device=dev)
    out = []
    for name, H, W, Cin, Cout in LAYERS:
        x = torch.randn(1, H, W, Cin, device=dev).bfloat16()
        w = (torch.randn(Cout, 3, 3, Cin, device=dev) / (9 * Cin) ** 0.5).bfloat16()
        dy = torch.randn(1, H, W, Cout, device=dev).bfloat16()
        bias = torch.zeros(Cout, device=dev)
        fl = 2.0 * H * W * Cin * Cout * 9
        out_w = torch.empty(Cout, 3, 3, Cin, device=dev)
        xc, wc, dyc = x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), dy.permute(0, 3, 1, 2)   # channels_last views
        lib = {"fwd": lambda: F.conv2d(xc, wc, None, padding=1),
               "dgrad": lambda: torch.nn.grad.conv2d_input(xc.shape, wc, dyc, padding=1),
               "wgrad": lambda: torch.nn.grad.conv2d_weight(xc, wc.shape, dyc, padding=1)}
        for what, fn in (("fwd", lambda: tc.conv3x3_nhwc(x, w, bias, relu=True)),
                         ("dgrad", lambda: tc.conv3x3_dgrad_nhwc(dy, w, mask_src=x)),
                         ("wgrad", lambda: tc.conv3x3_wgrad_nhwc(x, dy, out=out_w))):
            t = timeit(fn, flush)
            try:
                tl = timeit(lib[what], flush, warm=4)
            except Exception as e:                      # (a shape the library rejects)
                print("# %s %s: library call failed: %s" % (name, what, e), file=sys.stderr)
                tl = float("nan")
            out.append({"layer": name, "pass": what, "us": round(t * 1e6, 1),
                        "tflops": round(fl / t / 1e12, 1), "frac_bf16_peak": round(fl / t / 1e12 / pk, 3),
                        "lib_us": round(tl * 1e6, 1), "lib": "cuDNN bf16 channels_last",
                        "speedup_vs_lib": round(tl / t, 2)})
            print(json.dumps(out[-1]), flush=True)
    for name, M, N, K in GEMMS:
        a = torch.randn(M, K, device=dev).bfloat16()
        b = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        dyy = torch.randn(M, (N + 7) // 8 * 8, device=dev).bfloat16()
        fl = 2.0 * M * N * K
        bpad = torch.zeros((N + 7) // 8 * 8, K, device=dev, dtype=torch.bfloat16)
        bpad[:N] = b
        lib = {"fwd": lambda: torch.matmul(a, b.t()), "dgrad": lambda: torch.matmul(dyy, bpad),
               "wgrad": lambda: torch.matmul(dyy.t(), a)}
        for what, fn in (("fwd", lambda: tc.gemm_tn(a, b)),
                         ("dgrad", lambda: tc.gemm_nn(dyy, bpad)),
                         ("wgrad", lambda: tc.linear_wgrad(dyy, a))):
            t = timeit(fn, flush)
            tl = timeit(lib[what], flush, warm=4)
            out.append({"layer": name, "pass": what, "us": round(t * 1e6, 1),
                        "tflops": round(fl / t / 1e12, 1), "frac_bf16_peak": round(fl / t / 1e12 / pk, 3),
                        "lib_us": round(tl * 1e6, 1), "lib": "cuBLAS bf16 (bf16 output; ours writes fp32 weight gradients)",
                        "speedup_vs_lib": round(tl / t, 2)})
            print(json.dumps(out[-1]), flush=True)


if __name__ == "__main__":
    main()
