"""Phase boundaries of an iteration measured on the device (csrc/runtime.cu: scda_timestamp writes %globaltimer
in stream order), usable inside a captured graph.  Off unless SCDA_TIMESTAMPS=1: `mark(name)` is then one
1-thread kernel on the current stream; `read()` returns {name: microseconds since the first mark}."""
import os

import torch

ENABLED = os.environ.get("SCDA_TIMESTAMPS", "0") == "1"
_buf = None
_names = []


def mark(name):
    global _buf
    if not ENABLED or not torch.cuda.is_available():
        return
    from ._lib import check, load, stream_ptr
    if _buf is None:
        _buf = torch.zeros(256, dtype=torch.int64, device="cuda")
    if name not in _names:
        _names.append(name)
    slot = _names.index(name)
    check(load().scda_timestamp(_buf.data_ptr(), slot, stream_ptr(_buf.device)), "scda_timestamp")


def read():
    if _buf is None:
        return {}
    torch.cuda.synchronize()
    v = _buf.cpu().tolist()
    t0 = min(v[i] for i in range(len(_names)))
    return {n: (v[i] - t0) / 1e3 for i, n in enumerate(_names)}
