"""`pth_nms(dets, thresh)` — the reference's NMS entry point
(extensions/_nms/pth_nms.py:5-47, exported as `extensions.nms`).

Contract kept from the reference: `dets` is a float tensor [N, 5]
(x1, y1, x2, y2, score) already sorted by descending score, on any device (it
is moved to the GPU, :32); the result is a CPU LongTensor of kept indices in
ascending order (:47).  IoU uses the +1 box convention and a box is suppressed
iff IoU > thresh (strict).  Unlike gpu_nms (src/nms_cuda.c:17-67) the bitmask
never crosses PCIe: mask and scan both run on the device, and only the kept
indices are copied back.

`nms_device` is the form the on-device proposal pipeline uses: it returns
device tensors and never synchronises.
"""
import torch

from ..._lib import check, load, require_cuda, stream_ptr


def nms_device(dets, thresh, max_keep=0, workspace=None, n_dev=None):
    """dets: CUDA float32 [N, 5] contiguous, sorted by descending score.
    n_dev: optional int32 CUDA tensor [1]: only the first n_dev[0] rows are live.

    Returns (keep int64[N] device — first num_keep entries valid, num_keep int64[1] device).
    """
    require_cuda(dets)
    assert dets.dim() == 2 and dets.size(1) == 5 and dets.is_contiguous()
    assert dets.dtype == torch.float32
    n = dets.size(0)
    lib = load()
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=dets.device)
    num = torch.empty(1, dtype=torch.int64, device=dets.device)
    need = lib.scda_nms_workspace_bytes(n)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(max(need, 8) // 8, dtype=torch.int64, device=dets.device)
    with torch.cuda.device(dets.device):
        if n_dev is None:
            check(lib.scda_nms(n, dets.data_ptr(), float(thresh), int(max_keep), keep.data_ptr(),
                               num.data_ptr(), workspace.data_ptr(),
                               workspace.numel() * workspace.element_size(),
                               stream_ptr(dets.device)), "scda_nms")
        else:
            assert n_dev.is_cuda and n_dev.dtype == torch.int32 and n_dev.numel() == 1
            check(lib.scda_nms_dyn(n, n_dev.data_ptr(), dets.data_ptr(), float(thresh),
                                   int(max_keep), keep.data_ptr(), num.data_ptr(),
                                   workspace.data_ptr(),
                                   workspace.numel() * workspace.element_size(),
                                   stream_ptr(dets.device)), "scda_nms_dyn")
    return keep, num


def pth_nms(dets, thresh):
    if dets.numel() == 0:
        return torch.empty(0, dtype=torch.int64)
    dets = dets.cuda().contiguous().float()
    keep, num = nms_device(dets, thresh)
    n = int(num.item())
    return keep[:n].cpu().contiguous()
