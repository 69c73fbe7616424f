"""Operator micro-benchmarks on one GPU: libscda_b200 vs the reference's own kernels
(oracle/_ref/libscda_ref.so) under the same binding.  CUDA-event timing on the
launching stream, L2 flushed between iterations.  Prints one JSON line per op.
Usage: python scripts/opbench.py [--iters 20]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs  # noqa: E402
import _reflib  # noqa: E402
from scda_b200 import _lib  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, iters, flush):
    s = torch.cuda.current_stream()
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)           # 512 MB > 126 MB L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        fn()
        b.record(s)
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def burst(fn, n=50):
    """us per launch over n back-to-back launches (L2 warm): CUDA event timestamps tick every ~2 us on this
    part, too coarse for a single launch of the small operators"""
    s = torch.cuda.current_stream()
    fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(s)
    for _ in range(n):
        fn()
    b.record(s)
    b.synchronize()
    return a.elapsed_time(b) * 1e3 / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    ours = _lib.load()
    ref = _reflib.load() if _reflib.available() else None
    flush = torch.zeros(128 * 1024 * 1024, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def report(name, fn_ours, fn_ref, bytes_alg, extra=None):
        med, mn = timeit(fn_ours, a.iters, flush)
        row = {"op": name, "us": round(med, 2), "us_min": round(mn, 2)}
        if bytes_alg:
            row["GBs"] = round(bytes_alg / med / 1e3, 1)
            row["frac_hbm"] = round(bytes_alg / med / 1e3 / PEAK, 3)
        if med < 60:
            row["us_burst"] = round(burst(fn_ours), 2)
        if fn_ref is not None:
            rmed, _ = timeit(fn_ref, a.iters, flush)
            row["ref_us"] = round(rmed, 2)
            row["speedup_vs_ref_kernel"] = round(rmed / med, 2)
            if med < 60:
                row["ref_us_burst"] = round(burst(fn_ref), 2)
                row["speedup_burst"] = round(row["ref_us_burst"] / row["us_burst"], 2)
        if extra:
            row.update(extra)
        print(json.dumps(row), flush=True)

    # RoIPool at the model's operating point
    feat = torch.from_numpy(_inputs.features((1, 512, 32, 64), 0)).cuda()
    rois = torch.from_numpy(_inputs.rois_uniform(512, 1, img_w=1024, img_h=512)).cuda()
    out = torch.empty(512, 512, 7, 7, device="cuda")
    arg = torch.empty(512, 512, 7, 7, dtype=torch.int32, device="cuda")
    g = torch.randn_like(out)
    gi = torch.empty_like(feat)
    fb = feat.numel() * 4 + out.numel() * 8

    def pf(lib):
        return lambda: lib.ROIPoolForwardLaucher(feat.data_ptr(), 1 / 16., 512, 32, 64, 512, 7, 7,
                                                 rois.data_ptr(), out.data_ptr(), arg.data_ptr(), st)

    def pb(lib):
        return lambda: lib.ROIPoolBackwardLaucher(g.data_ptr(), 1 / 16., 1, 512, 32, 64, 512, 7, 7,
                                                  rois.data_ptr(), gi.data_ptr(), arg.data_ptr(), st)
    report("roi_pool_fwd[1x512x32x64,R512,7x7]", pf(ours), pf(ref) if ref else None, fb)
    report("roi_pool_bwd[1x512x32x64,R512,7x7]", pb(ours), pb(ref) if ref else None, fb)

    # RoIAlign config 1
    feat1 = torch.from_numpy(_inputs.features((1, 256, 64, 64), 0)).cuda()
    rois1 = torch.from_numpy(_inputs.rois_uniform(128, 1)).cuda()
    out1 = torch.empty(128, 256, 7, 7, device="cuda")
    g1 = torch.randn_like(out1)
    gi1 = torch.zeros_like(feat1)
    ab = feat1.numel() * 4 + out1.numel() * 4

    def af(lib):
        return lambda: lib.ROIAlignForwardLaucher(feat1.data_ptr(), 1 / 16., 128, 64, 64, 256, 7, 7,
                                                  rois1.data_ptr(), out1.data_ptr(), st)

    def abw(lib):
        return lambda: lib.ROIAlignBackwardLaucher(g1.data_ptr(), 1 / 16., 1, 128, 64, 64, 256, 7, 7,
                                                   rois1.data_ptr(), gi1.data_ptr(), st)
    report("roi_align_fwd[1x256x64x64,R128,7x7]", af(ours), af(ref) if ref else None, ab)
    report("roi_align_bwd[1x256x64x64,R128,7x7]", abw(ours), abw(ref) if ref else None, ab)

    # NMS sweep (config 5).  Reference = its kernel + D2H of the mask + host scan.
    import oracle
    for n in (1000, 2000, 6000, 12000, 30720, 100000):
        boxes = _inputs.nms_boxes(n, n)
        d = torch.from_numpy(boxes).cuda()
        keep = torch.empty(n, dtype=torch.int64, device="cuda")
        num = torch.zeros(1, dtype=torch.int64, device="cuda")
        wsb = ours.scda_nms_workspace_bytes(n)
        ws = torch.empty(wsb // 8 + 1, dtype=torch.int64, device="cuda")

        def f():
            ours.scda_nms(n, d.data_ptr(), 0.7, 0, keep.data_ptr(), num.data_ptr(), ws.data_ptr(), wsb, st)
        med, mn = timeit(f, a.iters, flush)
        row = {"op": "nms[N=%d,0.7]" % n, "us": round(med, 1), "us_min": round(mn, 1),
               "kept": int(num.item()), "pairs_per_s": round(n * (n - 1) / 2 / med * 1e6 / 1e9, 2)}
        if ref is not None and n <= 30720:
            import time
            cb = (n + 63) // 64
            mask = torch.empty(n, cb, dtype=torch.int64, device="cuda")
            ts = []
            for _ in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ref._nms(n, d.data_ptr(), mask.data_ptr(), 0.7)
                m = mask.cpu().numpy().view(np.uint64)
                oracle.nms_scan(m)
                ts.append((time.perf_counter() - t0) * 1e6)
            row["ref_pipeline_us(kernel+D2H+hostscan)"] = round(float(np.median(ts)), 1)
        print(json.dumps(row), flush=True)

    # IoU anchors x GT
    anchors = torch.from_numpy(_inputs.nms_boxes(30720, 1)[:, :4].copy()).cuda()
    for gk in (8, 32, 128):
        gts = torch.from_numpy(_inputs.gt_boxes(gk, 2)[:, :4].copy()).cuda()
        o = torch.empty(30720, gk, device="cuda")
        report("bbox_overlaps[30720x%d]" % gk,
               lambda: ours.scda_bbox_overlaps(30720, anchors.data_ptr(), gk, gts.data_ptr(), o.data_ptr(), st),
               None, 30720 * gk * 4 + 30720 * 16)
        report("IOUOverlap[30720x%d]" % gk,
               lambda: ours.IOUOverlap(anchors.data_ptr(), gts.data_ptr(), 4, 30720, gk, o.data_ptr(), st),
               (lambda: ref.IOUOverlap(anchors.data_ptr(), gts.data_ptr(), 4, 30720, gk, o.data_ptr(), st)) if ref else None,
               30720 * gk * 4 + 30720 * 16)

    # focal
    x, t = _inputs.focal_inputs(30720, 8, 0)
    xs, ts_ = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
    losses = torch.empty_like(xs)
    tot = torch.zeros(1, device="cuda")
    report("sigmoid_focal_fwd_sum[30720x8]",
           lambda: ours.scda_sigmoid_focal_loss_sum(xs.numel(), xs.data_ptr(), ts_.data_ptr(), 20.0, 2.0, 0.25, 8, None, tot.data_ptr(), st),
           (lambda: ref.SigmoidFocalLossForwardLaucher(xs.numel(), xs.data_ptr(), ts_.data_ptr(), 20.0, 2.0, 0.25, 8, losses.data_ptr(), st)) if ref else None,
           xs.numel() * 4 + ts_.numel() * 4)


if __name__ == "__main__":
    main()
