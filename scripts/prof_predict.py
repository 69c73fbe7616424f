import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import _inputs
from scda_b200.functions import predict_bbox as pb
from torch.profiler import profile, ProfilerActivity
cfg = dict(_inputs.load_cfg()["test_predict_bbox_cfg"], top_n=100)
r = np.random.RandomState(0); n = 300
rois = _inputs.rois_uniform(n, 20, img_w=1024, img_h=512, wh=(16, 300)); rois[:, 0] = 0
cls = torch.from_numpy(r.dirichlet(np.ones(9) * 0.3, n).astype(np.float32)).cuda()
loc = torch.from_numpy((r.standard_normal((n, 36)) * 0.5).astype(np.float32)).cuda()
rois = torch.from_numpy(rois).cuda(); info = np.array([[512, 1024, 1.0]], np.float32)
for _ in range(3): pb.compute_predicted_bboxes(rois, cls, loc, info, cfg)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): pb.compute_predicted_bboxes(rois, cls, loc, info, cfg)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
import time
t=time.perf_counter()
for _ in range(50): pb.compute_predicted_bboxes(rois, cls, loc, info, cfg)
torch.cuda.synchronize(); print("wall us per call", (time.perf_counter()-t)/50*1e6)
