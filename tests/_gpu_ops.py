"""Drive the Section A launcher ABI (include/scda_b200.h) of EITHER library —
libscda_b200.so or the reference's libscda_ref.so — with the same calls.
torch is used for device memory and the stream only."""
import numpy as np
import torch


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ok(st, what):
    assert st == 1, "%s returned %r" % (what, st)


def roi_pool_fwd(lib, feat, rois, ph, pw, scale):
    f, r = _dev(feat), _dev(rois)
    B, C, H, W = f.shape
    R = r.shape[0]
    out = torch.zeros(R, C, ph, pw, device="cuda")
    arg = torch.zeros(R, C, ph, pw, dtype=torch.int32, device="cuda")
    _ok(lib.ROIPoolForwardLaucher(f.data_ptr(), scale, R, H, W, C, ph, pw, r.data_ptr(),
                                  out.data_ptr(), arg.data_ptr(), _stream()), "ROIPoolForward")
    torch.cuda.synchronize()
    return out.cpu().numpy(), arg.cpu().numpy()


def roi_pool_bwd(lib, top_diff, rois, argmax, feat_shape, scale):
    g, r, a = _dev(top_diff), _dev(rois), _dev(argmax, torch.int32)
    B, C, H, W = feat_shape
    R, _, ph, pw = g.shape
    gi = torch.zeros(B, C, H, W, device="cuda")
    _ok(lib.ROIPoolBackwardLaucher(g.data_ptr(), scale, B, R, H, W, C, ph, pw, r.data_ptr(),
                                   gi.data_ptr(), a.data_ptr(), _stream()), "ROIPoolBackward")
    torch.cuda.synchronize()
    return gi.cpu().numpy()


def roi_align_fwd(lib, feat, rois, ah, aw, scale):
    f, r = _dev(feat), _dev(rois)
    B, C, H, W = f.shape
    R = r.shape[0]
    out = torch.zeros(R, C, ah, aw, device="cuda")
    _ok(lib.ROIAlignForwardLaucher(f.data_ptr(), scale, R, H, W, C, ah, aw, r.data_ptr(),
                                   out.data_ptr(), _stream()), "ROIAlignForward")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def roi_align_bwd(lib, top_diff, rois, feat_shape, scale):
    g, r = _dev(top_diff), _dev(rois)
    B, C, H, W = feat_shape
    R, _, ah, aw = g.shape
    gi = torch.zeros(B, C, H, W, device="cuda")
    _ok(lib.ROIAlignBackwardLaucher(g.data_ptr(), scale, B, R, H, W, C, ah, aw, r.data_ptr(),
                                    gi.data_ptr(), _stream()), "ROIAlignBackward")
    torch.cuda.synchronize()
    return gi.cpu().numpy()


def nms_mask(lib, boxes5, thresh):
    b = _dev(boxes5)
    n = b.shape[0]
    cb = (n + 63) // 64
    mask = torch.zeros(n, cb, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    lib._nms(n, b.data_ptr(), mask.data_ptr(), thresh)   # legacy default stream, like the reference
    torch.cuda.synchronize()
    return mask.cpu().numpy().view(np.uint64)


def iou_overlap(lib, b1, b2):
    a, b = _dev(b1[:, :4]), _dev(b2[:, :4])
    out = torch.zeros(a.shape[0], b.shape[0], device="cuda")
    _ok(lib.IOUOverlap(a.data_ptr(), b.data_ptr(), 4, a.shape[0], b.shape[0], out.data_ptr(),
                       _stream()), "IOUOverlap")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def sigmoid_focal(lib, logits, targets, weight_pos, gamma, alpha):
    x, t = _dev(logits), _dev(targets, torch.int32)
    M, K = x.shape
    losses = torch.zeros(M, K, device="cuda")
    dx = torch.zeros(M, K, device="cuda")
    _ok(lib.SigmoidFocalLossForwardLaucher(M * K, x.data_ptr(), t.data_ptr(), weight_pos, gamma,
                                           alpha, K, losses.data_ptr(), _stream()), "SigmoidFwd")
    _ok(lib.SigmoidFocalLossBackwardLaucher(M * K, x.data_ptr(), t.data_ptr(), dx.data_ptr(),
                                            weight_pos, gamma, alpha, K, _stream()), "SigmoidBwd")
    torch.cuda.synchronize()
    return losses.cpu().numpy(), dx.cpu().numpy()


def softmax_focal(lib, logits, targets, weight_pos, gamma, alpha):
    x, t = _dev(logits), _dev(targets, torch.int32)
    M, K = x.shape
    losses = torch.zeros(M, device="cuda")
    priors = torch.zeros(M, K, device="cuda")
    dx = torch.zeros(M, K, device="cuda")
    buff = torch.zeros(M, device="cuda")
    _ok(lib.SoftmaxFocalLossForwardLaucher(M * K, x.data_ptr(), t.data_ptr(), weight_pos, gamma,
                                           alpha, K, losses.data_ptr(), priors.data_ptr(),
                                           _stream()), "SoftmaxFwd")
    _ok(lib.SoftmaxFocalLossBackwardLaucher(M * K, x.data_ptr(), t.data_ptr(), dx.data_ptr(),
                                            weight_pos, gamma, alpha, K, priors.data_ptr(),
                                            buff.data_ptr(), _stream()), "SoftmaxBwd")
    torch.cuda.synchronize()
    return losses.cpu().numpy(), priors.cpu().numpy(), dx.cpu().numpy(), buff.cpu().numpy()
