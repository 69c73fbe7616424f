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
    names = sys.argv[1:] or list(SHAPES)
    dev = torch.device("cuda")
    data = {}
    for n in names:
        H, W, Cin, Cout = SHAPES[n]
        x = torch.randn(1, H, W, Cin, device=dev).bfloat16()
        w = (torch.randn(Cout, 3, 3, Cin, device=dev) / (9 * Cin) ** 0.5).bfloat16()
        b = torch.zeros(Cout, device=dev)
        data[n] = (x, w, b)
        for _ in range(2):
            tc.conv3x3_nhwc(x, w, b, relu=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for n in names:
        x, w, b = data[n]
        tc.conv3x3_nhwc(x, w, b, relu=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
