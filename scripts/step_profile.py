"""One training iteration between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/step_profile.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from scda_b200.engine import build_trainer  # noqa: E402

cfg = bench.load_cfg()
tr = build_trainer(cfg, seed=0)
image, target, gts, info = bench.synth_batch(0, pinned=False)
image, target, gts = image.cuda(), target.cuda(), gts.cuda()
for _ in range(3):
    tr.iteration(cfg, image, info, gts, target)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.iteration(cfg, image, info, gts, target)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one iteration")
