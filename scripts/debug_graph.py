import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs
from scda_b200.functions.proposal_target import proposal_targets_device, proposal_targets_tensor_ops
from scda_b200.functions.rpn_proposal import rpn_proposals_device
from scda_b200.functions._sampling import ArrayRng
cfg = _inputs.load_cfg()
g = np.load(os.path.join(ROOT, "tests/golden/host_plumbing.npz"))
gts = torch.from_numpy(g["gts"][0]).cuda()
cls, loc = _inputs.synth_rpn_outputs(0)
cls, loc = torch.from_numpy(cls).cuda(), torch.from_numpy(loc).cuda()
info = torch.from_numpy(g["image_info"])
pc, tc_ = cfg["train_rpn_proposal_cfg"], cfg["train_proposal_target_cfg"]
r = np.random.RandomState(5)
keys = [r.uniform(0, 1, 4096) for _ in range(3)]
dkeys = [torch.from_numpy(k).cuda() for k in keys]
class DevRng(object):
    def __init__(self): self.pos = 0
    def uniform(self, n, device):
        t = dkeys[self.pos][:n]; self.pos += 1; return t


def run():
    props = rpn_proposals_device(cls, loc, pc, info)
    boxes, n = props[0]
    return (boxes, n) + proposal_targets_device(boxes, n, gts, tc_, (512., 1024.), rng=DevRng())

def run_ref():
    props = rpn_proposals_device(cls, loc, pc, info)
    boxes, n = props[0]
    return (boxes, n) + proposal_targets_tensor_ops(boxes, n, gts, tc_, (512., 1024.), rng=DevRng())

eager = [t.clone() for t in run()]
ref = [t.clone() for t in run_ref()]
torch.cuda.synchronize()
for i, (a, b) in enumerate(zip(eager, ref)):
    print("eager vs tensor-ops", i, torch.equal(a, b), float((a.double() - b.double()).abs().max()))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        run()
torch.cuda.current_stream().wait_stream(s)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    outs = run()
for rep in range(3):
    gr.replay()
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(outs, eager)):
        print("replay", rep, i, torch.equal(a, b), bool(torch.isfinite(a.float()).all()))
