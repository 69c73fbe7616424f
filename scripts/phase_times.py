"""Device-side phase boundaries of the replayed iteration graph (no profiler): SCDA_TIMESTAMPS=1 python scripts/phase_times.py"""
import os, sys
os.environ["SCDA_TIMESTAMPS"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from scda_b200 import timestamps
from scda_b200.engine import build_trainer
cfg = bench.load_cfg()
tr = build_trainer(cfg, world_size=1, seed=0)
image, target, gts, info = bench.synth_batch(0, pinned=False)
image, target, gts = image.cuda(), target.cuda(), gts.cuda()
for _ in range(8):
    tr.iteration(cfg, image, info, gts, target)
acc = {}
N = 10
for _ in range(N):
    tr.iteration(cfg, image, info, gts, target)
    for k, v in timestamps.read().items():
        acc[k] = acc.get(k, 0.0) + v / N
for k, v in sorted(acc.items(), key=lambda kv: kv[1]):
    print("%8.1f us  %s" % (v, k))
