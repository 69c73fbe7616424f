#!/usr/bin/env python
"""Tile-plan sweep of the 3x3 convolution at the backbone shapes: per-tap kernel (gemm_tc.cu)
vs the halo kernel (conv_halo.cu) under each (N tile, sub-tiles) plan, forward and data
gradient.  CUDA events on the launch stream, L2 flushed between timed launches, median of 7.
One JSON line per (layer, pass, plan)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scda_b200 import tc  # noqa: E402
from scripts.convbench import LAYERS, peak, timeit  # noqa: E402

EXTRA = [("rpn3x3", 32, 64, 512, 512), ("dec_res", 64, 64, 128, 128), ("dec_up1", 128, 128, 128, 64)]
PLANS = [("per_tap", 0, 0, 0), ("halo_64x1", 1, 64, 1), ("halo_64x2", 1, 64, 2), ("halo_128x1", 1, 128, 1),
         ("halo_128x2", 1, 128, 2), ("pair_64", 1, 64, 0), ("pair_128", 1, 128, 0), ("pair_256", 1, 256, 0),
         ("halo_auto", 1, 0, 0)]


def main():
    dev = torch.device("cuda")
    pk = peak()
    flush = torch.zeros(64 * 1024 * 1024, device=dev)
    passes = sys.argv[1:] or ["fwd", "dgrad"]
    for name, H, W, Cin, Cout in LAYERS + EXTRA:
        NB = 4 if name.startswith("dec_") else 1
        x = torch.randn(NB, H, W, Cin, device=dev).bfloat16()
        w = (torch.randn(Cout, 3, 3, Cin, device=dev) / (9 * Cin) ** 0.5).bfloat16()
        dy = torch.randn(NB, H, W, Cout, device=dev).bfloat16()
        bias = torch.zeros(Cout, device=dev)
        fl = 2.0 * NB * H * W * Cin * Cout * 9
        fns = {"fwd": lambda: tc.conv3x3_nhwc(x, w, bias, relu=True),
               "dgrad": lambda: tc.conv3x3_dgrad_nhwc(dy, w, mask_src=x)}
        for what in passes:
            row = {}
            for plan, halo, bn, sub in PLANS:
                nout = Cout if what == "fwd" else Cin
                if bn >= 128 and nout % bn:
                    continue
                if plan == "pair_64" and what == "dgrad":
                    continue                   # (MN-major weight chunks are 64 columns: a pair needs N >= 128)
                tc.set_conv_pair(1 if plan.startswith("pair") else (-1 if plan == "halo_auto" else 0))
                tc.set_conv_plan(halo, bn, sub)
                t = timeit(fns[what], flush, iters=7, warm=2)
                row[plan] = round(t * 1e6, 1)
            best = min(row, key=row.get)
            print(json.dumps({"layer": name, "pass": what, "us": row, "best": best,
                              "best_tflops": round(fl / row[best] / 1e6, 1),
                              "best_frac_bf16_peak": round(fl / row[best] / 1e6 / pk, 3)}), flush=True)
    tc.set_conv_plan(1, 0, 0)
    tc.set_conv_pair(-1)


if __name__ == "__main__":
    main()
