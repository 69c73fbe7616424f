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
    dev = torch.device("cuda")
    pk = peak()
    flush = torch.zeros(64 * 1024 * 1024, device=dev)
    out = []
    for name, H, W, Cin, Cout in LAYERS:
        x = torch.randn(1, H, W, Cin, device=dev).bfloat16()
        w = (torch.randn(Cout, 3, 3, Cin, device=dev) / (9 * Cin) ** 0.5).bfloat16()
        dy = torch.randn(1, H, W, Cout, device=dev).bfloat16()
        bias = torch.zeros(Cout, device=dev)
        fl = 2.0 * H * W * Cin * Cout * 9
        for what, fn in (("fwd", lambda: tc.conv3x3_nhwc(x, w, bias, relu=True)),
                         ("dgrad", lambda: tc.conv3x3_dgrad_nhwc(dy, w, mask_src=x)),
                         ("wgrad", lambda: tc.conv3x3_wgrad_nhwc(x, dy))):
            t = timeit(fn, flush)
            out.append({"layer": name, "pass": what, "us": round(t * 1e6, 1),
                        "tflops": round(fl / t / 1e12, 1), "frac_bf16_peak": round(fl / t / 1e12 / pk, 3)})
            print(json.dumps(out[-1]), flush=True)
    for name, M, N, K in GEMMS:
        a = torch.randn(M, K, device=dev).bfloat16()
        b = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        dyy = torch.randn(M, (N + 7) // 8 * 8, device=dev).bfloat16()
        fl = 2.0 * M * N * K
        bpad = torch.zeros((N + 7) // 8 * 8, K, device=dev, dtype=torch.bfloat16)
        bpad[:N] = b
        for what, fn in (("fwd", lambda: tc.gemm_tn(a, b)),
                         ("dgrad", lambda: tc.gemm_nn(dyy, bpad)),
                         ("wgrad", lambda: tc.linear_wgrad(dyy, a))):
            t = timeit(fn, flush)
            out.append({"layer": name, "pass": what, "us": round(t * 1e6, 1),
                        "tflops": round(fl / t / 1e12, 1), "frac_bf16_peak": round(fl / t / 1e12 / pk, 3)})
            print(json.dumps(out[-1]), flush=True)


if __name__ == "__main__":
    main()
