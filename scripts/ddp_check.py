#!/usr/bin/env python
"""World-size-N sanity of the data-parallel iteration on real GPUs (run under torchrun): after a few
iterations (eager warm-up, capture, replays) every rank must hold bit-identical parameters in all
four networks (same all-reduced gradients -> same Adam step), and the parameters must have moved.
usage: torchrun --nproc-per-node 2 scripts/ddp_check.py [--no-graph-collectives] [--no-overlap]"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def main():
    from scda_b200.engine import build_trainer
    from scda_b200.utils.distributed_utils import broadcast_params
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = bench.load_cfg()
    tr = build_trainer(cfg, lr=1e-4, world_size=world, seed=rank,      # different init per rank ...
                       overlap="--no-overlap" not in sys.argv,
                       graph_collectives=False if "--no-graph-collectives" in sys.argv else
                       (True if "--graph-collectives" in sys.argv else None))
    for net in tr.nets():
        broadcast_params(net)                                          # ... made equal here
    for o in (tr.opt, tr.opt_dec, tr.opt_dis, tr.opt_dis_patch):       # shadows follow the masters
        if o.shadow is not None:
            o.shadow.copy_(o.flat)
    start = [o.flat.clone() for o in (tr.opt, tr.opt_dec, tr.opt_dis, tr.opt_dis_patch)]
    image, target, gts, info = bench.synth_batch(rank, pinned=False)
    image, target, gts = image.to(dev), target.to(dev), gts.to(dev)
    for it in range(5):
        out = tr.iteration(cfg, image, info, gts, target)
    torch.cuda.synchronize()
    ok = True
    for name, o, s0 in zip(("detector", "decoder", "dis", "dis_patch"),
                           (tr.opt, tr.opt_dec, tr.opt_dis, tr.opt_dis_patch), start):
        mine = o.flat
        ref = mine.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(mine, ref))
        moved = float((mine - s0).abs().max())
        flags = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("%-10s identical on all ranks: %s   max |update| %.3g" % (name, bool(flags.item()), moved))
        ok = ok and bool(flags.item()) and moved > 0
    if rank == 0:
        print("whole graph:", tr._whole_graph(), " loss", float(out["loss"]), " DDP_CHECK", "OK" if ok else "FAILED")
    sys.stdout.flush()
    tr.close()
    import threading
    threading.Timer(45.0, lambda: os._exit(0 if ok else 1)).start()
    dist.destroy_process_group()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
