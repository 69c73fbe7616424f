"""World-size-2 run of the data-parallel iteration on real GPUs (skipped on a one-GPU box):
scripts/ddp_check.py under torchrun — after eager warm-up, capture and replays, all four networks
must be bit-identical on every rank and must have moved (utils/distributed_utils.py:9-19 of the
reference: SUM all-reduce of the gradients, every rank applies the same step)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "scripts", "ddp_check.py")] + extra
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    return subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=420)


@pytest.mark.gpu
@pytest.mark.parametrize("extra,port", [(["--graph-collectives"], 29543),         # the default: NCCL captured in the one graph
                                        (["--no-graph-collectives"], 29541), (["--no-graph-collectives", "--no-overlap"], 29542)])
def test_two_ranks_hold_identical_parameters(cuda_lib, extra, port):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = _run(extra, port)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "DDP_CHECK OK" in r.stdout, r.stdout[-3000:]
    for net in ("detector", "decoder", "dis", "dis_patch"):
        assert any(l.startswith(net) and "identical on all ranks: True" in l for l in r.stdout.splitlines()), r.stdout[-3000:]
