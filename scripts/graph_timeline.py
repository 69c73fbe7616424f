#!/usr/bin/env python
"""Timeline of ONE replayed iteration graph (torch.profiler / CUPTI kernel records): span, sum of
kernel durations, time covered by at least one / at least two kernels, idle time between kernels,
per-stream totals.  Tells whether the step is bound by kernel execution or by launch gaps, and
whether the two captured streams really overlap.  usage: graph_timeline.py [--no-overlap]"""
import collections
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
    tr = build_trainer(cfg, world_size=1, seed=0, overlap="--no-overlap" not in sys.argv)
    image, target, gts, info = bench.synth_batch(0, pinned=False)
    image, target, gts = image.cuda(), target.cuda(), gts.cuda()
    for _ in range(5):
        tr.iteration(cfg, image, info, gts, target)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        tr.iteration(cfg, image, info, gts, target)
        torch.cuda.synchronize()
    import json
    trace = os.path.join(ROOT, "gpurun_out", "timeline_trace.json")
    os.makedirs(os.path.dirname(trace), exist_ok=True)
    prof.export_chrome_trace(trace)
    ks = []
    for ev in json.load(open(trace))["traceEvents"]:
        if ev.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in ev:
            ks.append((float(ev["ts"]), float(ev["ts"]) + float(ev["dur"]), ev.get("name", "?"),
                       ev.get("args", {}).get("stream", -1)))
    os.remove(trace)
    ks.sort()
    if not ks:
        print("no kernel records")
        return
    t0, t1 = ks[0][0], max(k[1] for k in ks)
    print("kernels %d  span %.1f us  sum of durations %.1f us" % (len(ks), t1 - t0, sum(k[1] - k[0] for k in ks)))
    # sweep
    evs = []
    for s, e, _, _ in ks:
        evs.append((s, 1))
        evs.append((e, -1))
    evs.sort()
    depth, last, cover = 0, evs[0][0], collections.Counter()
    for t, d in evs:
        cover[min(depth, 3)] += t - last
        depth += d
        last = t
    print("time with 0 kernels running %.1f us, exactly 1: %.1f, exactly 2: %.1f, 3+: %.1f" % (
        cover[0], cover[1], cover[2], cover[3]))
    per_stream = collections.Counter()
    cnt_stream = collections.Counter()
    for s, e, _, st in ks:
        per_stream[st] += e - s
        cnt_stream[st] += 1
    for st, t in per_stream.most_common():
        print("stream %s: %d kernels, %.1f us" % (st, cnt_stream[st], t))
    # duration histogram
    buckets = [(0, 2), (2, 4), (4, 8), (8, 16), (16, 32), (32, 64), (64, 1e9)]
    for lo, hi in buckets:
        sel = [k for k in ks if lo <= k[1] - k[0] < hi]
        print("kernels of %4g-%-6g us: %5d, total %.1f us" % (lo, hi, len(sel), sum(k[1] - k[0] for k in sel)))
    # the 25 longest idle gaps (no kernel running) and what follows them
    gaps = []
    end = ks[0][1]
    for s, e, name, _ in ks[1:]:
        if s > end:
            gaps.append((s - end, name[:60], s - t0))
        end = max(end, e)
    gaps.sort(reverse=True)
    print("idle gaps: %d, total %.1f us; longest:" % (len(gaps), sum(g[0] for g in gaps)))
    for g in gaps[:15]:
        print("   %7.1f us before %-60s at t=%.0f" % g)
    out = os.path.join(ROOT, "gpurun_out", "timeline.tsv")
    with open(out, "w") as f:
        for s, e, name, st in ks:
            f.write("%.1f\t%.1f\t%s\t%s\n" % (s - t0, e - s, st, name[:90]))


if __name__ == "__main__":
    main()
