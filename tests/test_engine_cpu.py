"""Host-side logic of the training engine that needs no GPU."""
import numpy as np
import torch


def test_device_crop_corners_equal_get_corner_from_center():
    """scda_crop_regions clamps int(c) - R/2 to [0, size - R] (engine.crop_corners states the rule);
    the reference's branchy get_corner_from_center (tools/faster_rcnn_train_val.py:411-438) yields
    the same windows."""
    from scda_b200.engine import crop_corners, get_corner_from_center
    R, W, H = 256, 1024, 512
    r = np.random.RandomState(0)
    centers = np.concatenate([
        np.stack([r.uniform(0, W, 200), r.uniform(0, H, 200)], 1),
        np.array([[0, 0], [W, H], [W - 1, H - 1], [127.9, 128.0], [128.0, 127.99], [896.0, 384.0],
                  [895.99, 383.5], [897, 385], [128, 128], [129, 129], [512.5, 256.5]])]).astype(np.float32)
    corners = get_corner_from_center(centers, R, W, H)
    x1, y1 = crop_corners(torch.from_numpy(centers), R, W, H)
    for k, (cx1, cy1, cx2, cy2) in enumerate(corners):
        assert cx2 - cx1 == R and cy2 - cy1 == R
        assert (int(x1[k]), int(y1[k])) == (cx1, cy1), (centers[k], (cx1, cy1))


def test_crops_device_has_no_cpu_path():
    import pytest
    from scda_b200.engine import crops_device
    with pytest.raises(RuntimeError):
        crops_device(torch.zeros(1, 3, 512, 1024), torch.zeros(4, 2), 256, 1024, 512)


def test_flat_layout_alignment():
    from scda_b200.utils.distributed_utils import flat_layout, flat_view
    ps = [torch.zeros(64, 3, 3, 3), torch.zeros(30), torch.zeros(60, 512, 1, 1), torch.zeros(9, 4096)]
    offs, total = flat_layout(ps, 64)
    assert all(o % 64 == 0 for o in offs) and total % 64 == 0
    flat = torch.arange(total, dtype=torch.float32)
    v = flat_view(flat, offs[0], ps[0])
    assert v.shape == (64, 3, 3, 3) and v.permute(0, 2, 3, 1).is_contiguous()
    assert flat_view(flat, offs[3], ps[3]).is_contiguous()
