#!/usr/bin/env python
"""Headline benchmark: images/sec of one SCDA training iteration (512x1024, bs 1 per GPU).

    python bench.py --gpus N --steps K --warmup W            # this build, N GPUs of one node
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path

Under `python -m torch.distributed.run --nproc-per-node N ...` each rank owns one GPU
(RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment); rank 0 prints ONE JSON line.

A "step" = one full iteration of the reference's train() on one synthetic source image, one
synthetic target image and 20 random ground-truth boxes per rank: detector forward on both
images, the four-phase discriminator / decoder / detector update, gradient all-reduce and
Adam on all four networks (BASELINE.json configs[3]; on one GPU this is configs[2] plus the
reconstruction networks).  value = images/s with the inputs already resident in HBM;
e2e = the same through the public step call with HOST (pinned) inputs copied in and the loss
read back every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

IMG_H, IMG_W, NUM_GT = 512, 1024, 20
METRIC, UNIT = "images/sec (512x1024, bs=1/GPU) fwd+bwd", "images/s"
WORKLOAD = "SCDA type-2 train iteration (cluster_num=4, threshold=128, recon_size=256): " \
           "vgg16 Faster R-CNN fwd(source)+fwd(target)+bwd + decoder/discriminators, Adam x4, " \
           "1x3x512x1024 source + target per GPU, 20 GT boxes"
# algorithmic work per image (SURVEY.md §8d / BASELINE.md §3)
CONV3x3_FWD_GFLOP_AT_MODEL_SHAPE = 38.65   # one 512->512 (or 64->64 @512x1024) 3x3 layer


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def load_cfg():
    from scda_b200 import synthetic
    return synthetic.load_cfg()


def synth_batch(rank, pinned):
    from scda_b200 import synthetic as _inputs
    r = np.random.RandomState(1000 + rank)
    mk = lambda: torch.from_numpy(r.standard_normal((1, 3, IMG_H, IMG_W)).astype(np.float32))
    image, target = mk(), mk()
    gts = torch.from_numpy(_inputs.gt_boxes(NUM_GT, rank, img_w=IMG_W, img_h=IMG_H)[None])
    info = torch.tensor([[IMG_H, IMG_W, 0.5]])
    if pinned:
        image, target, gts = image.pin_memory(), target.pin_memory(), gts.pin_memory()
    return image, target, gts, info


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def host_cores():
    """every core this process may run on (torchrun exports OMP_NUM_THREADS=1: ignored on purpose)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_run(steps, warmup, budget_s, threads=None):
    """The reference CPU-extension path (oracle/model_cpu.py) timed on ALL the host cores:
    torch's intra-op pool and the OpenMP loops of the C restatement are both set explicitly."""
    import oracle
    from oracle.model_cpu import CPUTrainer
    cores = int(threads) if threads else host_cores()
    torch.set_num_threads(cores)
    oracle.set_threads(cores)
    cfg = load_cfg()
    tr = CPUTrainer(cfg)
    image, target, gts, info = synth_batch(0, pinned=False)
    t_all = time.perf_counter()
    done_w = 0
    for _ in range(max(warmup, 1)):
        tr.iteration(image, info, gts, target)
        done_w += 1
        if time.perf_counter() - t_all > budget_s * 0.35:
            break
    times = []
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        tr.iteration(image, info, gts, target)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s and len(times) >= 1:
            break
    sec = float(np.mean(times))
    return {"value": 1.0 / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d warm-up + %d timed full iterations of the same workload (1 image each) on the "
                      "host, torch CPU fp32 + OpenMP C restatements of the reference's CUDA ops + the "
                      "reference's numpy/sklearn plumbing, %d threads (time-boxed to %ds)"
                      % (done_w, len(times), cores, budget_s),
            "ms_per_step": sec * 1e3, "steps_done": len(times), "warmup_done": done_w}


def reference_arm(args):
    """--impl reference: the CPU path on all host cores.  Under torchrun rank 0 alone runs it (the
    CPU path is ONE image stream whatever --gpus says; the line says so) and the other ranks exit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup, budget_s=args.cpu_budget)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": r["steps_done"], "warmup": r["warmup_done"],
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": "cpu x%d threads" % r["cores"],
                       "note": "the CPU path runs ONE image stream on all host cores regardless of --gpus: "
                               "only the N=1 ratio is like for like"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the 13 backbone launches, from the
# committed `ncu --set full` capture of the same kernels (scripts/prof_conv.py)
NCU_DRAM_BYTES_PER_LAUNCH = 24.15e6
NCU_TRAFFIC_SOURCE = ("profiles/r2_final_ncu_conv_fwd.txt: dram read + write of the 9 backbone shapes x multiplicity = 314.0 MB / 13 "
                      "launches (conv1_1 counted in its 64-channel halo form; round 1: 24.32 MB)")

VGG_CONVS = [(3, 64, 512, 1024), (64, 64, 512, 1024), (64, 128, 256, 512), (128, 128, 256, 512),
             (128, 256, 128, 256), (256, 256, 128, 256), (256, 256, 128, 256), (256, 512, 64, 128),
             (512, 512, 64, 128), (512, 512, 64, 128), (512, 512, 32, 64), (512, 512, 32, 64),
             (512, 512, 32, 64)]     # (Cin, Cout, H, W) of the 13 backbone convolutions at 512 x 1024


def roofline_probe(dev, x3=False):
    """Dominant hand-written kernel = the tcgen05 halo convolution (conv_halo_kernel):
    the 13 backbone launches of one image timed back to back with CUDA events on the launch
    stream (L2 flushed before each pass).  Algorithmic work = 2 * H * W * Cin * Cout * 9 per
    layer (conv1_1 counted with its 3 real input channels) = 320.71 GFLOP per image
    (SURVEY.md section 8d); achieved = that / the time of the 13 launches.
    x3: the fp32-parity mode — the same kernel on split operands (three bf16 MMAs per product,
    csrc/x3_ops.cu), fp32 output; the peak it is held against is bf16_tflops / 3."""
    from scda_b200 import tc
    hbm, tf, tf_sus, src = peaks()
    xs, ws, bs = [], [], []
    mult = 3 if x3 else 1
    from scda_b200 import tc_detector
    direct = tc_detector.FIRST_DIRECT and not x3       # conv1_1 reads the fp32 NCHW image (csrc/conv_first.cu)
    for cin, cout, h, w in VGG_CONVS:
        if cin < 64 and direct:
            xs.append(torch.randn(1, cin, h, w, device=dev))
            ws.append((torch.randn(cout, 3, 3, cin, device=dev) / (9 * cin) ** 0.5).bfloat16())
        else:
            cp = max(cin, 64) * mult
            xs.append(torch.randn(1, h, w, cp, device=dev).bfloat16())
            ws.append((torch.randn(cout, 3, 3, cp, device=dev) / (9 * cp) ** 0.5).bfloat16())
        bs.append(torch.zeros(cout, device=dev))
    flops = sum(2.0 * h * w * cin * cout * 9 for cin, cout, h, w in VGG_CONVS)
    flush = torch.zeros(64 * 1024 * 1024, device=dev)
    ts = []
    for i in range(8):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for x, w, bias in zip(xs, ws, bs):
            if x.dtype == torch.float32:
                tc.conv3x3_first_nchw(x, w, bias, relu=True)
            else:
                tc.conv3x3_nhwc(x, w, bias, relu=True, out_dtype=torch.float32 if x3 else torch.bfloat16)
        b.record()
        b.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b) * 1e-3)
    t = float(np.mean(ts))
    peak = tf / mult
    return {"bound": "tensor", "achieved": flops / t / 1e12, "peak": peak, "unit": "TFLOP/s",
            "frac": flops / t / 1e12 / peak, "traffic": None if x3 else NCU_DRAM_BYTES_PER_LAUNCH,
            "traffic_source": None if x3 else NCU_TRAFFIC_SOURCE,
            "kernel": "conv_halo_kernel (tcgen05 + TMA 3x3 convolution, input halo staged once for the nine taps, "
                      "+ bias + ReLU), 13 backbone launches"
                      + (" (conv1_1 = conv_first_kernel: im2col in shared memory from the fp32 NCHW image)" if direct else "")
                      + (" on hi/lo split operands (3 bf16 MMAs per fp32 product), fp32 output" if x3 else ""),
            "peak_source": src + " (MEASURED_PEAKS.json bf16_tflops, burst: kernels timed alone)"
                           + (" / 3: three tensor-core products per algorithmic product" if x3 else ""),
            "algorithmic_flops_per_launch": flops / len(VGG_CONVS),
            "avg_launch_us": t / len(VGG_CONVS) * 1e6}


def aux_ops(dev):
    """BASELINE.json's second metric, "RoIAlign+NMS us/image": the reference's RoIAlign forward at config 1
    (1x256x64x64 features, 128 RoIs, 7x7), the RoI max-pool the model really uses (NHWC bf16, 512 RoIs on the
    1x512x32x64 map) and NMS(0.7) over 12 000 score-sorted synthetic proposals (config 5 generator), each through
    the C ABI, CUDA events on the launch stream, L2 flushed, median of 10."""
    from scda_b200 import _lib, tc
    from scda_b200 import synthetic as _inputs
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.zeros(64 * 1024 * 1024, device=dev)

    def timeit(fn):
        ts = []
        for i in range(13):
            flush.add_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b) * 1e3)
        return float(np.median(ts))
    feat1 = torch.from_numpy(_inputs.features((1, 256, 64, 64), 0)).to(dev)
    rois1 = torch.from_numpy(_inputs.rois_uniform(128, 1)).to(dev)
    out1 = torch.empty(128, 256, 7, 7, device=dev)
    t_align = timeit(lambda: lib.ROIAlignForwardLaucher(feat1.data_ptr(), 1 / 16., 128, 64, 64, 256, 7, 7,
                                                        rois1.data_ptr(), out1.data_ptr(), st))
    featn = torch.randn(1, 32, 64, 512, device=dev).bfloat16()
    rois = torch.from_numpy(_inputs.rois_uniform(512, 1, img_w=IMG_W, img_h=IMG_H)).to(dev)
    t_pool = timeit(lambda: tc.roi_pool_nhwc(featn, rois, 7, 7, 1 / 16.))
    n = 12000
    boxes = torch.from_numpy(_inputs.nms_boxes(n, n)).to(dev)
    keep = torch.empty(n, dtype=torch.int64, device=dev)
    num = torch.zeros(1, dtype=torch.int64, device=dev)
    wsb = lib.scda_nms_workspace_bytes(n)
    ws = torch.empty(wsb // 8 + 1, dtype=torch.int64, device=dev)
    t_nms = timeit(lambda: lib.scda_nms(n, boxes.data_ptr(), 0.7, 0, keep.data_ptr(), num.data_ptr(), ws.data_ptr(),
                                        wsb, st))
    return {"roi_align_fwd_us": round(t_align, 1), "roi_pool_nhwc_fwd_us": round(t_pool, 1),
            "nms_12k_us": round(t_nms, 1), "nms_12k_kept": int(num.item()),
            "roi_align_plus_nms_us_per_image": round(t_align + t_nms, 1),
            "what": "RoIAlign fwd 1x256x64x64/128 RoIs + NMS(0.7) of 12000 sorted boxes, on-device scan included"}


def our_arm(args):
    import torch.distributed as dist
    from scda_b200 import _lib
    from scda_b200.engine import build_trainer
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path in scda_b200)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    from scda_b200 import gan_ops, tc
    tc.set_precision(args.precision)
    cfg = load_cfg()
    tr = build_trainer(cfg, world_size=world, seed=0, use_graphs=not args.no_graphs,
                       overlap=not args.no_overlap,
                       graph_collectives=False if args.no_graph_collectives else (True if args.graph_collectives else None),
                       force_cut=args.force_cut)
    if world > 1:
        from scda_b200.utils.distributed_utils import broadcast_params
        for net in tr.nets():
            broadcast_params(net)
    h_image, h_target, h_gts, info = synth_batch(rank, pinned=True)
    d_image, d_target, d_gts = h_image.to(dev), h_target.to(dev), h_gts.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return tr.iteration(cfg, d_image, info, d_gts, d_target)

    def step_e2e():
        # host (pinned) inputs in, loss out: the call a user of the engine makes
        out = tr.iteration(cfg, h_image, info, h_gts, h_target)
        return float(out['loss'].item())           # D2H read of the step's result

    # our kernels launched per iteration, counted on the first (eager) iteration; the
    # replayed CUDA graph contains exactly these launches
    n0 = _lib.LAUNCHES
    step_resident()
    per_iter = _lib.LAUNCHES - n0

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step_resident()
    b.record()
    barrier()
    launches = per_iter * args.steps
    ms = a.elapsed_time(b)
    clocks = sampler.summary()

    step_e2e()
    barrier()
    a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a2.record()
    # the training loop a user writes: every step's inputs go host -> device inside the timed region (the NEXT
    # step's copy is started on the engine's copy stream before this step's loss is read back, as a data
    # loader would), and every step's loss is read back
    tr.prefetch(h_image, h_gts, h_target)
    for i in range(args.steps):
        out = tr.iteration(cfg, h_image, info, h_gts, h_target)
        if i + 1 < args.steps:
            tr.prefetch(h_image, h_gts, h_target)
        last = float(out['loss'].item())
    b2.record()
    barrier()
    ms2 = a2.elapsed_time(b2)

    whole_graph, overlap_on = tr._whole_graph(), tr.overlap
    if world > 1:
        t = torch.tensor([ms, ms2], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms2 = float(t[0]), float(t[1])
    parity = None
    fallbacks_main = dict(gan_ops.LIBRARY_CALLS)        # library (cuDNN / torch) layers met by the headline run
    if rank == 0 and world == 1 and args.precision == "bf16" and not args.no_parity_line:
        # the same iteration in the fp32-parity precision mode (tc.set_precision('bf16x3')), timed in
        # this process right after the headline run: a second trainer, 3 warm-up + 10 timed steps
        del tr
        torch.cuda.empty_cache()
        tc.set_precision("bf16x3")
        tr3 = build_trainer(cfg, world_size=1, seed=0, use_graphs=not args.no_graphs, overlap=not args.no_overlap)
        for _ in range(4):
            tr3.iteration(cfg, d_image, info, d_gts, d_target)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(10):
            tr3.iteration(cfg, d_image, info, d_gts, d_target)
        p1.record()
        torch.cuda.synchronize()
        pms = p0.elapsed_time(p1) / 10
        roof3 = roofline_probe(dev, x3=True)
        parity = {"dtype": "bf16x3 (fp32 activations; operands split into bf16 hi + lo, 3 MMAs per product, "
                           "~2^-16 per product; the reference's arithmetic is fp32)",
                  "value": 1e3 / pms, "unit": UNIT, "ms_per_step": pms, "steps": 10, "warmup": 4,
                  "roofline": {k: roof3[k] for k in ("bound", "achieved", "peak", "unit", "frac", "peak_source")},
                  "parity": "tests/test_iteration_parity_gpu.py, profiles/r2_iteration_parity.txt",
                  # (this mode keeps the two discriminators on cuDNN fp32: they have no split-operand form)
                  "library_fallbacks": {k: v - fallbacks_main.get(k, 0) for k, v in gan_ops.LIBRARY_CALLS.items()
                                        if v - fallbacks_main.get(k, 0) > 0}}
        del tr3
        tc.set_precision("bf16")
        tr = None
    if rank == 0:
        roof = roofline_probe(dev, x3=args.precision == "bf16x3")
        line = {"metric": METRIC, "value": world * args.steps / (ms / 1e3), "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": {"workload": WORKLOAD, "parallelism": "dp%d" % world,
                           "execution": ("eager" if args.no_graphs else
                                         "one CUDA graph per iteration" + (", NCCL all-reduces captured" if world > 1 else "")
                                         if whole_graph else
                                         "eleven CUDA graphs per iteration on three streams, cut at the gradient all-reduces "
                                         "(NCCL issued between them)")
                           + (", detector backward + Adam overlapped with the reconstruction/discriminator "
                              "updates on a second stream, target-image branch beside the source branch"
                              if overlap_on else ""),
                           "l2": "per-step working set (547 MB of fp32 weights + activations) exceeds the "
                                 "126 MB L2; no explicit flush"},
                "clocks": clocks,
                "e2e": {"value": world * args.steps / (ms2 / 1e3), "unit": UNIT,
                        "h2d_bytes_per_step": int(h_image.numel() * 4 + h_target.numel() * 4 + h_gts.numel() * 4),
                        "d2h_bytes_per_step": 4, "last_loss": last,
                        "input_copy": "pinned host -> device on the engine's copy stream (SCDATrainer.prefetch), "
                                      "started before the previous step's loss is read back; every copy is "
                                      "inside the timed region"},
                "gpu_launches": launches, "roofline": roof,
                "library_fallbacks": fallbacks_main}
        if parity is not None:
            line["parity_mode"] = parity
        if world == 1:
            line["aux"] = aux_ops(dev)
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(2, 1, budget_s=args.cpu_budget_inline)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        # captured NCCL kernels pin their communicator: drop the graphs first; and never let a stuck
        # teardown hold the launcher (the line is already printed)
        if tr is not None:
            tr.close()
        threading.Timer(45.0, lambda: os._exit(0)).start()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=int, default=240, help="seconds for the --impl reference arm")
    ap.add_argument("--cpu-budget-inline", type=int, default=45)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="eager execution (for kernel profilers)")
    ap.add_argument("--no-overlap", action="store_true", help="single stream (no detector/GAN overlap)")
    ap.add_argument("--force-cut", action="store_true", help="1 GPU: replay the cut (world > 1) graph plan")
    ap.add_argument("--precision", "--dtype", dest="precision", default="bf16", choices=["bf16", "bf16x3"],
                    help="bf16: throughput mode; bf16x3: fp32-parity mode (3 bf16 MMAs per product)")
    ap.add_argument("--no-parity-line", action="store_true", help="skip the bf16x3 sub-measurement")
    ap.add_argument("--no-graph-collectives", action="store_true",
                    help="world > 1: cut the graph at the all-reduces instead of capturing NCCL")
    ap.add_argument("--graph-collectives", action="store_true",
                    help="world > 1: capture the NCCL all-reduces inside the one iteration graph")
    ap.add_argument("--config", type=int, choices=[2, 3, 5], default=None,
                    help="BASELINE.json configs[1] / [2] / [4] instead of the headline configs[3]: forward-only "
                         "per-stage us / detector-only train step / NMS + IoU sweep (one JSON line, 1 GPU)")
    args = ap.parse_args()
    if args.config is not None:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import configbench
        torch.cuda.set_device(0)
        print(json.dumps(configbench.run(args.config)), flush=True)
        return
    if args.impl == "reference":
        reference_arm(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    our_arm(args)


if __name__ == "__main__":
    main()
