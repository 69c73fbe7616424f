#!/usr/bin/env python
"""Launch a few representative tcgen05 convolution shapes once inside a cudaProfiler range, for
`ncu --set full --profile-from-start off` (see /opt/skills/guides/B200_PROFILING.md)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scda_b200 import tc  # noqa: E402

# the 9 distinct shapes of the 13 backbone convolutions (H, W, Cin, Cout); conv1_1 has its 3 input
# channels zero-padded to 64.  Multiplicity in the backbone: conv3_2 x2, conv4_2 x2, conv5_x x3.
SHAPES = {"conv1_1p": (512, 1024, 64, 64), "conv1_2": (512, 1024, 64, 64), "conv2_1": (256, 512, 64, 128),
          "conv2_2": (256, 512, 128, 128), "conv3_1": (128, 256, 128, 256), "conv3_2": (128, 256, 256, 256),
          "conv4_1": (64, 128, 256, 512), "conv4_2": (64, 128, 512, 512), "conv5_x": (32, 64, 512, 512)}


def main():
    """usage: prof_conv.py [--pass=fwd|dgrad|wgrad] [layer ...]"""
    what = "fwd"
    names = []
    for a in sys.argv[1:]:
        if a.startswith("--pass="):
            what = a.split("=", 1)[1]
        else:
            names.append(a)
    names = names or list(SHAPES)
    dev = torch.device("cuda")
    data = {}

    def run(x, w, b, dy, out):
        if what == "fwd":
            tc.conv3x3_nhwc(x, w, b, relu=True)
        elif what == "dgrad":
            tc.conv3x3_dgrad_nhwc(dy, w, mask_src=x)
        else:
            tc.conv3x3_wgrad_nhwc(x, dy, out=out)        # the weight-gradient kernel + reduce_slabs

    for n in names:
        H, W, Cin, Cout = SHAPES[n]
        x = torch.randn(1, H, W, Cin, device=dev).bfloat16()
        w = (torch.randn(Cout, 3, 3, Cin, device=dev) / (9 * Cin) ** 0.5).bfloat16()
        b = torch.zeros(Cout, device=dev)
        dy = torch.randn(1, H, W, Cout, device=dev).bfloat16()
        out = torch.empty(Cout, 3, 3, Cin, device=dev)
        data[n] = (x, w, b, dy, out)
        for _ in range(2):
            run(*data[n])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for n in names:
        run(*data[n])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
