"""Datasets of the reference driver (mirror datasets/example_dataset.py:20-145, target_dataset.py:23-71,
example_loader.py:8-56) with the pixel work on the GPU.

Reference per image, on one host core: PIL open -> resize -> (flip) -> ToTensor -> Normalize -> pad / stack in
the collate function -> pageable H2D of fp32.  Here the dataset hands over the DECODED uint8 image (HWC) and
the already-transformed boxes (host arithmetic on a few numbers, the reference's floor / ceil / mirror rules);
`prepare_image` uploads the uint8 pixels and ONE kernel (scda_image_prepare) writes the normalised fp32 NCHW
network input: resize + mirror + / 255 + (x - 0.5) / 0.5.

Meta file format (example_dataset.py:40-68): per image
    # <index>
    <relative path>
    <channels>
    <height>
    <width>
    <flag>
    <number of ignore boxes>      followed by that many `x1 y1 x2 y2` lines
    <number of ground-truth boxes> followed by that many `label x1 y1 x2 y2` lines
"""
import os

import numpy as np
import torch

from .._lib import check, load, stream_ptr

MEAN = (0.5, 0.5, 0.5)
STD = (0.5, 0.5, 0.5)


def parse_meta_file(list_file):
    """-> list of [img_name, height, width, gt [G, 4], labels [G], ignores [I, 4]] (example_dataset.py:40-68)"""
    with open(list_file) as f:
        lines = f.readlines()
    metas, i = [], 0
    while i < len(lines):
        if not lines[i].strip():
            i += 1
            continue
        img_name = lines[i + 1].rstrip()
        h, w = float(lines[i + 3]), float(lines[i + 4])
        n_ig = int(lines[i + 6])
        i += 7
        ig = [[float(v) for v in lines[i + j].split()[:4]] for j in range(n_ig)] or [[0, 0, 0, 0]]
        i += n_ig
        n_gt = int(lines[i])
        i += 1
        gt, labels = [], []
        for j in range(n_gt):
            sp = lines[i + j].split()
            gt.append([float(sp[1]), float(sp[2]), float(sp[3]), float(sp[4])])
            labels.append(int(sp[0]))
        i += n_gt
        metas.append([img_name, h, w, np.array(gt, dtype=np.float64).reshape(-1, 4), np.array(labels),
                      np.array(ig, dtype=np.float64)])
    return metas


class ExampleTransform(object):
    """scale jitter + mirror of the reference (example_dataset.py:107-145) as BOX arithmetic and a pixel recipe:
    __call__(w, h, bbox, ignores) -> (new_w, new_h, scale, flip, new_bbox, new_ignores)."""

    def __init__(self, sizes, max_size, flip=False, rng=None):
        sizes = sizes if isinstance(sizes, (list, tuple)) else [sizes]
        self.scale_min, self.scale_max, self.max_size, self.flip = min(sizes), max(sizes), max_size, flip
        self.rng = rng or np.random

    @staticmethod
    def _scale_boxes(b, scale):
        b = np.array(b, dtype=np.float64)
        if b.shape[0] > 0:
            b[:, 0], b[:, 1] = np.floor(b[:, 0] * scale), np.floor(b[:, 1] * scale)
            b[:, 2], b[:, 3] = np.ceil(b[:, 2] * scale), np.ceil(b[:, 3] * scale)
        return b

    def __call__(self, w, h, bbox, ignores):
        size = self.rng.randint(self.scale_min, self.scale_max + 1)
        scale = min(size / min(w, h), self.max_size / max(w, h))
        new_w, new_h = int(w * scale), int(h * scale)
        nb, ni = self._scale_boxes(bbox, scale), self._scale_boxes(ignores, scale)
        flip = bool(self.flip and self.rng.random() < 0.5)
        if flip:
            if nb.shape[0] > 0:
                nb[:, 0], nb[:, 2] = new_w - nb[:, 2], new_w - nb[:, 0].copy()
            if ni.shape[0] > 0:
                ni[:, 0], ni[:, 2] = new_w - ni[:, 2], new_w - ni[:, 0].copy()
        return new_w, new_h, scale, flip, nb, ni


def _decode(path):
    from PIL import Image
    img = Image.open(path)
    if img.mode != 'RGB':
        img = img.convert('RGB')
    return np.array(img, dtype=np.uint8)            # a writable copy (torch.as_tensor warns on read-only views)


def prepare_image(pixels_hwc_u8, new_h, new_w, flip=False, mode="nearest", device="cuda", out=None):
    """uint8 HWC (numpy or tensor, host or device) -> fp32 [1, 3, new_h, new_w] on `device`: resize, mirror,
    ToTensor, Normalize(0.5, 0.5) in one kernel.  mode 'nearest' = Pillow < 7's Image.resize default (the
    reference's era), 'bilinear' for modern Pillow-like smoothing."""
    src = torch.as_tensor(pixels_hwc_u8)
    assert src.dtype == torch.uint8 and src.dim() == 3 and src.shape[2] == 3
    if not src.is_cuda:
        src = (src.pin_memory() if torch.cuda.is_available() else src).to(device, non_blocking=True)
    src = src.contiguous()
    if out is None:
        out = torch.empty(1, 3, new_h, new_w, dtype=torch.float32, device=src.device)
    import ctypes
    mean = (ctypes.c_float * 3)(*MEAN)
    std = (ctypes.c_float * 3)(*STD)
    with torch.cuda.device(src.device):
        check(load().scda_image_prepare(src.data_ptr(), src.shape[0], src.shape[1], out.data_ptr(), new_h, new_w,
                                        {"nearest": 0, "bilinear": 1}[mode], 1 if flip else 0,
                                        ctypes.cast(mean, ctypes.c_void_p), ctypes.cast(std, ctypes.c_void_p),
                                        stream_ptr(src.device)), "scda_image_prepare")
    return out


class ExampleDataset(torch.utils.data.Dataset):
    """source-domain dataset: returns (uint8 HWC pixels, recipe) — the pixels are finished on the device by
    `collate` / `prepare_image`.  recipe = (new_h, new_w, scale, flip, boxes [G, 5] = x1,y1,x2,y2,label,
    ignores, filename)."""

    def __init__(self, root_dir, list_file, transform_fn, normalize_fn=None):
        self.root_dir, self.transform_fn = root_dir, transform_fn
        self.metas = parse_meta_file(list_file)
        self.num = len(self.metas)
        self.aspect_ratios = [float(m[1]) / m[2] for m in self.metas]

    def __len__(self):
        return self.num

    def __getitem__(self, idx):
        name, h, w, bbox, labels, ignores = self.metas[idx]
        filename = os.path.join(self.root_dir, name)
        pixels = _decode(filename)
        assert pixels.shape[1] == w and pixels.shape[0] == h
        new_w, new_h, scale, flip, nb, ni = self.transform_fn(int(w), int(h), bbox, ignores)
        boxes = np.hstack([nb, labels.astype(np.float64)[:, None]]).astype(np.float32)
        return pixels, (new_h, new_w, scale, flip, boxes, ni.astype(np.float32), filename)


class TargetDataset(torch.utils.data.Dataset):
    """target-domain images resized to (new_h, new_w), no labels (target_dataset.py:23-71)"""

    def __init__(self, root_dir, list_file, normalize_fn=None, new_w=1024, new_h=512):
        self.root_dir, self.new_w, self.new_h = root_dir, new_w, new_h
        with open(list_file) as f:
            self.metas = [x.strip() for x in f.readlines() if x.strip()]
        self.num = len(self.metas)

    def __len__(self):
        return self.num

    def __getitem__(self, idx):
        return _decode(os.path.join(self.root_dir, self.metas[idx])), (self.new_h, self.new_w)


def collate(batch, device="cuda", mode="nearest"):
    """the reference's collate (example_loader.py:14-56): images padded with zeros at the right / bottom to the
    largest of the batch, boxes / ignores padded with zero rows -> (images [B,3,H,W] on `device`, image_info
    [B,3] = (h, w, scale), gts [B,G,5], ignores [B,I,4], filenames)"""
    H = max(r[0] for _, r in batch)
    W = max(r[1] for _, r in batch)
    G = max(max(r[4].shape[0] for _, r in batch), 1)
    Ig = max(max(r[5].shape[0] for _, r in batch), 1)
    images = torch.zeros(len(batch), 3, H, W, dtype=torch.float32, device=device)
    gts = np.zeros((len(batch), G, 5), np.float32)
    igs = np.zeros((len(batch), Ig, 4), np.float32)
    info, names = [], []
    for b, (pixels, (nh, nw, scale, flip, boxes, ignores, fn)) in enumerate(batch):
        if nh == H and nw == W:
            prepare_image(pixels, nh, nw, flip, mode, device, out=images[b:b + 1])
        else:
            images[b, :, :nh, :nw] = prepare_image(pixels, nh, nw, flip, mode, device)[0]
        gts[b, :boxes.shape[0]] = boxes
        igs[b, :ignores.shape[0]] = ignores
        info.append([nh, nw, scale])
        names.append(fn)
    return images, torch.tensor(info, dtype=torch.float32), torch.from_numpy(gts), torch.from_numpy(igs), names


class SyntheticPairs(object):
    """seeded synthetic (source, target) pairs of the benchmark shape for runs without a dataset on disk:
    yields (image [1,3,H,W], image_info [1,3], gts [1,G,5], target [1,3,H,W]) host tensors."""

    def __init__(self, n, new_h=512, new_w=1024, num_gt=20, seed=0, num_classes=9):
        from .. import synthetic
        self.n, self.h, self.w, self.g, self.seed, self.nc, self.syn = n, new_h, new_w, num_gt, seed, num_classes, synthetic

    def __len__(self):
        return self.n

    def __iter__(self):
        for i in range(self.n):
            r = np.random.RandomState(self.seed * 100003 + i)
            mk = lambda: torch.from_numpy(r.standard_normal((1, 3, self.h, self.w)).astype(np.float32))
            gts = torch.from_numpy(self.syn.gt_boxes(self.g, self.seed * 100003 + i, img_w=self.w, img_h=self.h,
                                                     num_classes=self.nc)[None])
            yield mk(), torch.tensor([[self.h, self.w, 1.0]]), gts, mk()
