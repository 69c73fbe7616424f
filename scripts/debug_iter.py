"""debug: one eager iteration at 256x512, print every output and tap summary"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs
from scda_b200.engine import build_trainer
H, W = 256, 512
cfg = _inputs.load_cfg()
tr = build_trainer(cfg, lr=1e-4, new_w=W, new_h=H, world_size=1, seed=0, use_graphs=('--graphs' in sys.argv), overlap=('--no-overlap' not in sys.argv))
r = np.random.RandomState(0)
img = torch.from_numpy(r.standard_normal((1, 3, H, W)).astype(np.float32)).cuda()
tgt = torch.from_numpy(r.standard_normal((1, 3, H, W)).astype(np.float32)).cuda()
gts = torch.from_numpy(_inputs.gt_boxes(12, 0, img_w=W, img_h=H)[None]).cuda()
info = torch.tensor([[H, W, 0.5]])
if '--taps' in sys.argv:
    tr.taps = {}
for it in range(3):
    torch.manual_seed(it)
    out = tr.iteration(cfg, img, info, gts, tgt)
    torch.cuda.synchronize()
    print("iteration", it, {k: round(float(v), 4) for k, v in out.items()})
    d = getattr(tr.model, "_dbg", None)
    if d:
        for i, nm in enumerate(("rois", "labels", "loc_t", "loc_w")):
            e, l, o = d['early'][i], d['late'][i], d['orig'][i]
            print("   ", nm, "early finite", bool(torch.isfinite(e.float()).all()), "late==early", torch.equal(e, l),
                  "orig==early", torch.equal(e, o), "min/max", float(e.float().min()), float(e.float().max()))
        print("    pred finite", [bool(torch.isfinite(t).all()) for t in d['pred']])
def summ(name, t):
    if torch.is_tensor(t):
        tf = t.float()
        print("tap %-16s shape=%s finite=%s min=%.4g max=%.4g" % (name, tuple(t.shape), bool(torch.isfinite(tf).all()), float(tf.min()) if tf.numel() else 0, float(tf.max()) if tf.numel() else 0))
    elif isinstance(t, (list, tuple)):
        for i, x in enumerate(t):
            summ("%s[%d]" % (name, i), x)
    elif isinstance(t, dict):
        for kk, x in t.items():
            summ("%s.%s" % (name, kk), x)
    else:
        print("tap", name, t)
for k, v in (tr.taps or {}).items():
    summ(k, v)
