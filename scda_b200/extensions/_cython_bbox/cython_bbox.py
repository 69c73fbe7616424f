"""`bbox_overlaps(boxes, query_boxes)` with the reference's Cython module's
signature and arithmetic (extensions/_cython_bbox/cython_bbox.pyx:32-73; called
through utils/bbox_helper.py:8-9 for anchor and proposal targets): float32
[N, 4] x [K, 4] -> float32 [N, K]; no +1, zero unless both overlaps > 0.
Computed on the GPU, bit-identical to the host code."""
import numpy as np
import torch

from ..._lib import check, load, stream_ptr


def bbox_overlaps_device(boxes, query_boxes):
    """CUDA float32 [N, 4] x [K, 4] -> CUDA float32 [N, K]; no host sync."""
    assert boxes.is_cuda and query_boxes.is_cuda
    assert boxes.dtype == torch.float32 and query_boxes.dtype == torch.float32
    assert boxes.is_contiguous() and query_boxes.is_contiguous()
    assert boxes.dim() == 2 and boxes.size(1) == 4
    assert query_boxes.dim() == 2 and query_boxes.size(1) == 4
    out = torch.empty(boxes.size(0), query_boxes.size(0), dtype=torch.float32,
                      device=boxes.device)
    with torch.cuda.device(boxes.device):
        check(load().scda_bbox_overlaps(boxes.size(0), boxes.data_ptr(), query_boxes.size(0),
                                        query_boxes.data_ptr(), out.data_ptr(),
                                        stream_ptr(boxes.device)), "scda_bbox_overlaps")
    return out


def bbox_overlaps(boxes, query_boxes):
    if boxes.dtype != np.float32 or query_boxes.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float32'")  # Cython's typed-buffer check
    if boxes.ndim != 2 or query_boxes.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2)")
    b = torch.from_numpy(np.ascontiguousarray(boxes[:, :4])).cuda()
    q = torch.from_numpy(np.ascontiguousarray(query_boxes[:, :4])).cuda()
    return bbox_overlaps_device(b, q).cpu().numpy()
