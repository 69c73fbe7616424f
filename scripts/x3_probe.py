"""fp32-parity mode probe: per-parameter gradient error of the x3 detector stages and of the plain fp32
torch graph, BOTH against an fp64 torch graph (how much of the difference is fp32's own noise)."""
import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scda_b200 import synthetic as _inputs, tc
from scda_b200.models.faster_rcnn.vgg_adver_expansion_cluster import vgg16

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
cfg = _inputs.load_cfg()
torch.manual_seed(0)
model = vgg16(cfg=cfg["shared"]).cuda()
with torch.no_grad():
    for n, p in model.named_parameters():
        if n.endswith("bias"):
            p.normal_(0, 0.05)
model.eval()
g = torch.Generator(device="cuda").manual_seed(1)
img = torch.randn(1, 3, 128, 256, device="cuda", generator=g)
rois = torch.from_numpy(_inputs.rois_uniform(64, 3, img_w=256, img_h=128, wh=(16, 128))).cuda()
m64 = copy.deepcopy(model).double()

def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp(min=1e-30))

def run(mode):
    m = m64 if mode == "fp64" else model
    m.zero_grad()
    m._fp32_graph = mode != "x3"
    x = img.double() if mode == "fp64" else img
    feat = m.feature_extractor(x)
    cls, loc = m.rpn(feat)
    # (loss on the RPN outputs only: the RoIPool extension has no fp64 form)
    m._fp32_graph = False
    dt = cls.dtype
    w1 = torch.linspace(-1, 1, cls.numel(), device="cuda", dtype=dt).view_as(cls)
    w2 = torch.linspace(1, -1, loc.numel(), device="cuda", dtype=dt).view_as(loc)
    loss = (cls * w1).sum() + (loc * w2).sum()
    loss.backward()
    return {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}, feat.detach()

tc.set_precision("bf16x3")
try:
    g64, f64 = run("fp64")
except Exception as e:
    print("fp64 run failed:", e); raise
g32, f32 = run("fp32")
gx3, fx3 = run("x3")
print("feat: fp32 vs fp64 %.2e   x3 vs fp64 %.2e" % (rel(f32, f64), rel(fx3.permute(0, 3, 1, 2), f64)))
for n in g64:
    print("%-28s fp32-vs-fp64 %.2e   x3-vs-fp64 %.2e   x3-vs-fp32 %.2e" % (n, rel(g32[n], g64[n]), rel(gx3[n], g64[n]), rel(gx3[n], g32[n])))
