"""Anchor generation (mirrors utils/anchor_helper.py of the reference).

`get_anchors_over_plane(featmap_h, featmap_w, anchor_ratios, anchor_scales, anchor_stride)`
returns the same float64 [K*A, 4] array as the reference (cell-major, anchor-minor;
ratio-major, scale-minor inside a cell).  Quirk kept for parity: the reference never uses
`anchor_ratios` (utils/anchor_helper.py:4-11 returns before reading it) and always
enumerates aspect ratios (0.5, 1, 2).

Anchors depend only on the feature-map size and the config, so they are computed once and
cached, on the host and (anchors_device) as a float64 CUDA tensor.
"""
import functools

import numpy as np

_ASPECT = (0.5, 1.0, 2.0)


@functools.lru_cache(maxsize=None)
def _grid(scales, stride):
    size = float(stride) * float(stride)
    ctr = 0.5 * (stride - 1)
    rows = []
    for r in _ASPECT:
        w = np.round(np.sqrt(size / r))
        h = np.round(w * r)
        for s in scales:
            ww, hh = w * float(s), h * float(s)
            rows.append((ctr - 0.5 * (ww - 1), ctr - 0.5 * (hh - 1),
                         ctr + 0.5 * (ww - 1), ctr + 0.5 * (hh - 1)))
    out = np.array(rows, dtype=np.float64)
    out.setflags(write=False)
    return out


def get_anchors_over_grid(ratios, scales, stride):
    del ratios  # ignored by the reference as well
    return _grid(tuple(float(np.float64(s) * stride / stride) for s in scales), int(stride)).copy()


@functools.lru_cache(maxsize=None)
def _plane(fh, fw, scales, stride):
    grid = _grid(scales, stride)
    xs = np.arange(fw, dtype=np.float64) * stride
    ys = np.arange(fh, dtype=np.float64) * stride
    shift = np.stack([np.tile(xs, fh), np.repeat(ys, fw)] * 2, axis=1)      # [K, 4]
    out = (shift[:, None, :] + grid[None, :, :]).reshape(-1, 4)
    out.setflags(write=False)
    return out


def get_anchors_over_plane(featmap_h, featmap_w, anchor_ratios, anchor_scales, anchor_stride):
    del anchor_ratios
    return _plane(int(featmap_h), int(featmap_w),
                  tuple(float(np.float64(s) * anchor_stride / anchor_stride) for s in anchor_scales),
                  int(anchor_stride)).copy()


_DEVICE_CACHE = {}


def anchors_device(featmap_h, featmap_w, anchor_ratios, anchor_scales, anchor_stride, device):
    """float64 CUDA tensor [K*A, 4], cached per (shape, config, device)."""
    import torch
    key = (int(featmap_h), int(featmap_w), tuple(float(s) for s in anchor_scales), int(anchor_stride),
           str(device))
    t = _DEVICE_CACHE.get(key)
    if t is None:
        t = torch.from_numpy(get_anchors_over_plane(featmap_h, featmap_w, anchor_ratios,
                                                    anchor_scales, anchor_stride)).to(device)
        _DEVICE_CACHE[key] = t
    return t
