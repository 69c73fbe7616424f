"""Data-parallel plumbing (mirrors utils/distributed_utils.py of the reference).

Same three entry points — `dist_init`, `broadcast_params`, `average_gradients` — with the
reference's semantics (SUM all-reduce of every gradient, loss pre-divided by world size by
the caller, tools/faster_rcnn_train_val.py:604,736).  Differences in mechanism:
  * rendezvous comes from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT), falling back to the reference's SLURM variables;
  * `average_gradients` reduces ONE flat buffer per network instead of one NCCL call per
    parameter tensor (40 calls / 547 MB for the detector in the reference,
    utils/distributed_utils.py:9-13): see FlatGradBucket.
"""
import os

import torch
import torch.distributed as dist


def dist_init(port=None, backend="nccl"):
    """Returns (rank, world_size).  One process per GPU."""
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        local = int(os.environ.get("LOCAL_RANK", rank))
    elif "SLURM_PROCID" in os.environ:            # the reference's launch path (:26-45)
        rank, world = int(os.environ["SLURM_PROCID"]), int(os.environ["SLURM_NTASKS"])
        local = rank % max(torch.cuda.device_count(), 1)
        node_list = os.environ["SLURM_NODELIST"]
        if "[" in node_list:
            beg = node_list.find("[")
            ends = [p for p in (node_list.find("-", beg), node_list.find(",", beg)) if p >= 0]
            node_list = node_list[:min(ends) if ends else 1000].replace("[", "")
        os.environ.setdefault("MASTER_ADDR", node_list[8:].replace("-", "."))
        os.environ["RANK"], os.environ["WORLD_SIZE"] = str(rank), str(world)
    else:
        rank, world, local = 0, 1, 0
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["RANK"], os.environ["WORLD_SIZE"] = "0", "1"
    if port is not None:
        os.environ.setdefault("MASTER_PORT", str(port))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return dist.get_rank(), dist.get_world_size()


def broadcast_params(model):
    """Rank 0's state to everyone (:15-19), coalesced into one broadcast per dtype."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    by_dtype = {}
    for p in model.state_dict().values():
        by_dtype.setdefault((p.dtype, p.device), []).append(p)
    for tensors in by_dtype.values():
        flat = torch.cat([t.reshape(-1) for t in tensors])
        dist.broadcast(flat, 0)
        off = 0
        for t in tensors:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()


def flat_layout(params, align=64):
    """offset of each parameter in a flat buffer; every segment starts on an `align`-element
    boundary (16-byte aligned fp32 AND bf16 views for TMA) -> (offsets, total length)"""
    offs, off = [], 0
    for p in params:
        offs.append(off)
        off += (p.numel() + align - 1) // align * align
    return offs, off


def flat_view(flat, off, p):
    """view of flat[off : off + p.numel()] with p's shape; 4-D weights are laid out
    [O][kh][kw][I] (channels_last), the order the NHWC convolution kernels read."""
    seg = flat[off:off + p.numel()]
    if p.dim() == 4:
        o, i, kh, kw = p.shape
        return seg.view(o, kh, kw, i).permute(0, 3, 1, 2)
    return seg.view(p.shape)


# FlatGradBucket.zero() does not zero a direct-written gradient of at least this many elements (its writer
# overwrites it); smaller ones ARE zeroed there, so a sink that finds such a gradient marked fresh has nothing to clear
DIRECT_SKIP_NUMEL = 1 << 20


class FlatGradBucket(object):
    """All gradients of one network as views into one contiguous fp32 buffer, so the
    data-parallel exchange is a single all-reduce and the optimiser can sweep one array."""

    def __init__(self, params, offsets=None, total=None, channels_last=False):
        self.params = [p for p in params if p.requires_grad]
        if offsets is None:
            offsets, total = flat_layout(self.params, 1)
        self.offsets, self.channels_last = offsets, channels_last
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self._views = [self._view(p, off) for p, off in zip(self.params, offsets)]
        for p, v in zip(self.params, self._views):
            p.grad = v

    def _view(self, p, off):
        if self.channels_last:
            return flat_view(self.flat, off, p)
        return self.flat[off:off + p.numel()].view_as(p)

    def zero(self):
        """zero everything except gradients that a direct writer will overwrite
        (`p._scda_direct_grad` and large): those are only marked fresh."""
        skip = []
        for p, off in zip(self.params, self.offsets):
            if getattr(p, "_scda_direct_grad", False):
                p._scda_grad_fresh = True
                if p.numel() >= DIRECT_SKIP_NUMEL:
                    skip.append((off, off + p.numel()))
        if not skip:
            self.flat.zero_()
            return
        pos = 0
        for a, b in skip:
            if a > pos:
                self.flat[pos:a].zero_()
            pos = b
        if pos < self.flat.numel():
            self.flat[pos:].zero_()

    def settle(self, lo=0, hi=None):
        """a gradient still marked fresh was never written this step: it is zero
        (lo, hi: only the parameters whose segment starts inside [lo, hi) of the flat buffer)"""
        for p, v, off in zip(self.params, self._views, self.offsets):
            if off < lo or (hi is not None and off >= hi):
                continue
            if getattr(p, "_scda_grad_fresh", False):
                if p.numel() >= DIRECT_SKIP_NUMEL:
                    v.zero_()
                p._scda_grad_fresh = False

    def rebind(self):
        """Re-point p.grad at the flat buffer if something replaced it."""
        for p, view in zip(self.params, self._views):
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view

    def all_reduce(self, async_op=False, lo=0, hi=None, group=None):
        """SUM all-reduce of the flat buffer, or of its slice [lo, hi) (a gradient bucket), on
        `group` (None = the default process group)"""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            buf = self.flat if (lo == 0 and hi is None) else self.flat[lo:hi]
            return dist.all_reduce(buf, async_op=async_op, group=group)
        return None


def average_gradients(model):
    """SUM all-reduce of the model's gradients (:9-13).  If the model carries a
    FlatGradBucket (attribute `_scda_bucket`) that single buffer is reduced; otherwise the
    gradients are coalesced into one message."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    bucket = getattr(model, "_scda_bucket", None)
    if bucket is not None:
        bucket.rebind()
        bucket.all_reduce()
        return
    grads = [p.grad.data for p in model.parameters() if p.requires_grad and p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
