"""Data-parallel plumbing at world size 2 on the gloo backend (CPU): `broadcast_params`,
`average_gradients` (SUM semantics, utils/distributed_utils.py:9-19 of the reference) and the
flat gradient bucket whose single all-reduce replaces the per-tensor collectives."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _net(seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 4, 1), nn.Flatten(),
                         nn.Linear(4 * 6 * 6, 5))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from scda_b200.utils.distributed_utils import (FlatGradBucket, average_gradients, broadcast_params,
                                                   dist_init, flat_layout)
    r, w = dist_init(backend="gloo")
    assert (r, w) == (rank, world)
    net = _net(100 + rank)                       # different weights per rank before the broadcast
    broadcast_params(net)
    ref = _net(100)                              # what rank 0 started with
    for a, b in zip(net.state_dict().values(), ref.state_dict().values()):
        assert torch.equal(a, b)

    x = torch.randn(2, 3, 6, 6, generator=torch.Generator().manual_seed(rank))
    # (a) no bucket: gradients are summed over ranks, tensor by tensor semantics preserved
    (net(x).sum() / world).backward()
    local = [p.grad.clone() for p in net.parameters()]
    average_gradients(net)
    summed = [p.grad.clone() for p in net.parameters()]
    gathered = [None] * world
    dist.all_gather_object(gathered, [g.numpy() for g in local])
    for i, g in enumerate(summed):
        want = sum(torch.from_numpy(gathered[k][i]) for k in range(world))
        assert torch.allclose(g, want, rtol=1e-6, atol=1e-7)

    # (b) flat bucket with the aligned, channels_last layout: ONE all-reduce, same result
    net.zero_grad()
    params = list(net.parameters())
    offs, total = flat_layout(params, 64)
    bucket = FlatGradBucket(params, offs, total, channels_last=True)
    net._scda_bucket = bucket
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, bucket._views))
    (net(x).sum() / world).backward()
    bucket.rebind()
    for p, want in zip(params, local):
        assert torch.allclose(p.grad, want, rtol=1e-6, atol=1e-7)
    average_gradients(net)                       # takes the bucket path
    for p, want in zip(params, summed):
        assert torch.allclose(p.grad, want, rtol=1e-6, atol=1e-7)
    assert float(bucket.flat[offs[1] - 1]) == 0.0 or offs[1] == params[0].numel()   # padding stays zero
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_flat_bucket_and_broadcast_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out.get() == "ok"
