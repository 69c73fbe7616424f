"""One set of uniform key vectors drives both sides of a sampling parity test:
the device pipeline (scda_b200.functions._sampling.ArrayRng replays them in call order)
and the oracle (a `choice` function with np.random.choice's meaning)."""
import numpy as np


class KeyedChoice(object):
    """choice(n, size, replace=False) = the `size` ranks with the smallest keys[tag][:n], in
    key order; choice(n, size, replace=True) = floor(keys[tag][:size] * n)."""
    wants_tag = True

    def __init__(self, keys_by_tag):
        self.keys = keys_by_tag

    def __call__(self, n, size=None, replace=True, tag=None):
        k = np.asarray(self.keys[tag], dtype=np.float64)
        if replace:
            return np.floor(k[:size] * n).astype(np.int64)
        return np.argsort(k[:n], kind="stable")[:size]
