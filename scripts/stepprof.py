#!/usr/bin/env python
"""torch.profiler view of one training iteration: which aten op / autograd node owns the GPU
time (kernel names alone do not say which layer a cuDNN / cutlass kernel belongs to)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def main():
    from torch.profiler import ProfilerActivity, profile
    from scda_b200.engine import build_trainer
    torch.cuda.set_device(0)
    cfg = bench.load_cfg()
    tr = build_trainer(cfg, world_size=1, seed=0, use_graphs=False, overlap=False)
    image, target, gts, info = bench.synth_batch(0, pinned=False)
    image, target, gts = image.cuda(), target.cuda(), gts.cuda()
    for _ in range(3):
        tr.iteration(cfg, image, info, gts, target)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        tr.iteration(cfg, image, info, gts, target)
        torch.cuda.synchronize()
    print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=90,
                                                             max_name_column_width=60, max_shapes_column_width=70))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))


if __name__ == "__main__":
    main()
