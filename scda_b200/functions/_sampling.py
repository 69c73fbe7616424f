"""Random subset selection on the device without host round trips.

The reference draws with `np.random.choice(n, size, replace=False)` on the host, where n
is the (data dependent) number of candidates (functions/anchor_target.py:82-94,
functions/proposal_target.py:99-111,149-155).  Here every draw is expressed with a vector
of uniform keys: candidate j (j-th in ascending index order) owns keys[j]; "choose k of n
without replacement" = the k candidates with the smallest keys, in key order.  All shapes
are static (masks + counts kept as device scalars), so the whole target stage can be
captured in a CUDA graph.

An `rng` object supplies the keys: `TorchRng` (torch.rand on the device) in production,
`ArrayRng` (prescribed arrays) in parity tests — the oracle is driven by a `choice`
function built from the same arrays (tests/_sampling_adapter.py).
"""
import torch


_CONST = {}


def const_tensor(values, dtype, device):
    """small constant tensor built once per (values, dtype, device): creating it from a
    Python list on every call is an unpinned H2D copy, which a CUDA-graph capture forbids"""
    key = (tuple(float(v) for v in values), dtype, str(device))
    t = _CONST.get(key)
    if t is None:
        t = torch.tensor(list(values), dtype=dtype, device=device)
        _CONST[key] = t
    return t


def scalar_i64(v, device):
    """0-dim int64 device tensor from a Python int (a fill kernel, not an H2D copy) or from
    a device scalar (no host read): both forms are legal inside a CUDA-graph capture"""
    if torch.is_tensor(v):
        return v.to(device=device, dtype=torch.int64).reshape(())
    return torch.full((), int(v), dtype=torch.int64, device=device)


class TorchRng(object):
    def __init__(self, generator=None):
        self.generator = generator

    def uniform(self, n, device):
        return torch.rand(n, device=device, dtype=torch.float64, generator=self.generator)


class ArrayRng(object):
    """Replays prescribed key vectors, in call order."""

    def __init__(self, arrays):
        self.arrays = list(arrays)
        self.pos = 0

    def uniform(self, n, device):
        import numpy as np
        a = np.asarray(self.arrays[self.pos], dtype=np.float64)
        self.pos += 1
        assert a.shape[0] >= n, "prescribed key vector too short"
        return torch.from_numpy(a[:n].copy()).to(device)


def ranks_of(mask):
    """rank[i] = number of set entries before i (valid where mask), count as 0-dim tensor."""
    c = torch.cumsum(mask.to(torch.int64), 0)
    return c - 1, c[-1] if mask.numel() else torch.zeros((), dtype=torch.int64, device=mask.device)


def members_by_rank(mask, rank):
    """out[j] = index of the j-th set entry (garbage for j >= count)."""
    n = mask.numel()
    out = torch.zeros(n + 1, dtype=torch.int64, device=mask.device)
    out.scatter_(0, torch.where(mask, rank, torch.full_like(rank, n)),
                 torch.arange(n, device=mask.device))
    return out[:n]


def key_order(keys, count):
    """Ranks sorted by ascending key among the first `count` ranks (others pushed last)."""
    n = keys.numel()
    k = torch.where(torch.arange(n, device=keys.device) < count, keys,
                    torch.full_like(keys, float("inf")))
    return torch.sort(k, stable=True)[1]


def choose(mask, want, rng, kmax):
    """Pick min(want, count) members of `mask`.

    want: python int or 0-dim int64 tensor.  Returns (idx int64 [kmax], n_sel 0-dim):
    if count > want the `want` members with the smallest keys in key order (what
    `cand[np.random.choice(count, want, replace=False)]` yields for the key-driven choice),
    else all members in ascending order.  Entries beyond n_sel are unspecified.
    """
    dev = mask.device
    n = mask.numel()
    rank, count = ranks_of(mask)
    members = members_by_rank(mask, rank)
    want_t = scalar_i64(want, dev)
    keys = rng.uniform(n, dev)          # drawn unconditionally: keeps the call sequence static
    order = key_order(keys, count)
    ar = torch.arange(n, device=dev)
    sel_rank = torch.where(count > want_t, order, ar)
    idx = members[sel_rank.clamp(max=n - 1)]
    if kmax <= n:
        idx = idx[:kmax]
    else:
        idx = torch.cat([idx, idx.new_zeros(kmax - n)])
    return idx, torch.minimum(count, want_t)


def drop(mask, n_keep, rng):
    """Remove (count - n_keep) members of `mask`, the ones with the smallest keys (what
    `cand[np.random.choice(count, count - n_keep, replace=False)]` removes); no-op when
    count <= n_keep.  Returns (new mask, new count)."""
    dev = mask.device
    n = mask.numel()
    rank, count = ranks_of(mask)
    n_keep_t = scalar_i64(n_keep, dev)
    n_remove = (count - n_keep_t).clamp(min=0)
    keys = rng.uniform(n, dev)
    order = key_order(keys, count)
    ar = torch.arange(n, device=dev)
    removed_rank = torch.zeros(n, dtype=torch.bool, device=dev)
    removed_rank.scatter_(0, order, ar < n_remove)
    gone = mask & removed_rank[rank.clamp(min=0)]
    return mask & ~gone, torch.minimum(count, n_keep_t)
