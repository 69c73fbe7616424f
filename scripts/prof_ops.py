"""Launch each operator a few times so ncu can capture it:
   ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 2 -o gpurun_out/x python scripts/prof_ops.py <op>"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs  # noqa: E402
from scda_b200 import _lib  # noqa: E402

op = sys.argv[1] if len(sys.argv) > 1 else "all"
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.zeros(128 * 1024 * 1024, device="cuda")

if op in ("roi_pool", "all"):
    feat = torch.from_numpy(_inputs.features((1, 512, 32, 64), 0)).cuda()
    rois = torch.from_numpy(_inputs.rois_uniform(512, 1, img_w=1024, img_h=512)).cuda()
    out = torch.empty(512, 512, 7, 7, device="cuda")
    arg = torch.empty(512, 512, 7, 7, dtype=torch.int32, device="cuda")
    g = torch.randn_like(out)
    gi = torch.empty_like(feat)
    for _ in range(3):
        flush.add_(1.0)
        lib.ROIPoolForwardLaucher(feat.data_ptr(), 1 / 16., 512, 32, 64, 512, 7, 7, rois.data_ptr(),
                                  out.data_ptr(), arg.data_ptr(), st)
        flush.add_(1.0)
        lib.ROIPoolBackwardLaucher(g.data_ptr(), 1 / 16., 1, 512, 32, 64, 512, 7, 7, rois.data_ptr(),
                                   gi.data_ptr(), arg.data_ptr(), st)
if op in ("roi_align", "all"):
    feat1 = torch.from_numpy(_inputs.features((1, 256, 64, 64), 0)).cuda()
    rois1 = torch.from_numpy(_inputs.rois_uniform(128, 1)).cuda()
    out1 = torch.empty(128, 256, 7, 7, device="cuda")
    g1 = torch.randn_like(out1)
    gi1 = torch.zeros_like(feat1)
    for _ in range(3):
        flush.add_(1.0)
        lib.ROIAlignForwardLaucher(feat1.data_ptr(), 1 / 16., 128, 64, 64, 256, 7, 7, rois1.data_ptr(),
                                   out1.data_ptr(), st)
        lib.ROIAlignBackwardLaucher(g1.data_ptr(), 1 / 16., 1, 128, 64, 64, 256, 7, 7, rois1.data_ptr(),
                                    gi1.data_ptr(), st)
if op in ("nms", "all"):
    n = 12000
    d = torch.from_numpy(_inputs.nms_boxes(n, n)).cuda()
    keep = torch.empty(n, dtype=torch.int64, device="cuda")
    num = torch.zeros(1, dtype=torch.int64, device="cuda")
    wsb = lib.scda_nms_workspace_bytes(n)
    ws = torch.empty(wsb // 8 + 1, dtype=torch.int64, device="cuda")
    for _ in range(3):
        flush.add_(1.0)
        lib.scda_nms(n, d.data_ptr(), 0.7, 0, keep.data_ptr(), num.data_ptr(), ws.data_ptr(), wsb, st)
torch.cuda.synchronize()
print("done", op)
