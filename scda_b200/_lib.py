"""ctypes binding of libscda_b200.so (the C ABI declared in include/scda_b200.h).

There is no fallback of any kind: if the shared object is missing or a symbol
is absent, importing the operators raises.  Tensors cross the boundary as raw
device pointers (`tensor.data_ptr()`), sizes and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libscda_b200.so")

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_z = C.c_size_t

# name -> (restype, argtypes); mirrors include/scda_b200.h one to one
SIGNATURES = {
    "ROIPoolForwardLaucher": (_i, [_p, _f, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "ROIPoolBackwardLaucher": (_i, [_p, _f, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "ROIAlignForwardLaucher": (_i, [_p, _f, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "ROIAlignBackwardLaucher": (_i, [_p, _f, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "_nms": (None, [_i, _p, _p, _f]),
    "IOUOverlap": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "SigmoidFocalLossForwardLaucher": (_i, [_i, _p, _p, _f, _f, _f, _i, _p, _p]),
    "SigmoidFocalLossBackwardLaucher": (_i, [_i, _p, _p, _p, _f, _f, _f, _i, _p]),
    "SoftmaxFocalLossForwardLaucher": (_i, [_i, _p, _p, _f, _f, _f, _i, _p, _p, _p]),
    "SoftmaxFocalLossBackwardLaucher": (_i, [_i, _p, _p, _p, _f, _f, _f, _i, _p, _p, _p]),
    "scda_abi_version": (_i, []),
    "scda_nms_workspace_bytes": (_z, [_i]),
    "scda_nms": (_i, [_i, _p, _f, _i, _p, _p, _p, _z, _p]),
    "scda_nms_dyn": (_i, [_i, _p, _p, _f, _i, _p, _p, _p, _z, _p]),
    "scda_nms_mask": (_i, [_i, _p, _p, _f, _p]),
    "scda_nms_groups": (_i, [_i, _i, _p, _p, _f, _p, _p, _p]),
    "scda_predict_prepare": (_i, [_i, _i, _p, _i, _p, _p, _i, _p, _p, C.c_double, C.c_double, _f, _p, _p, _p]),
    "scda_predict_topn": (_i, [_i, _i, _p, _p, _p, _f, _i, _p, _p, _p]),
    "scda_bbox_overlaps": (_i, [_i, _p, _i, _p, _p, _p]),
    "scda_sigmoid_focal_loss_sum": (_i, [_i, _p, _p, _f, _f, _f, _i, _p, _p, _p]),
    "scda_softmax_focal_loss_sum": (_i, [_i, _p, _p, _f, _f, _f, _i, _p, _p, _p, _p]),
    "scda_gemm_bf16_tn": (_i, [_i, _i, _i, _p, C.c_longlong, _p, C.c_longlong, _p, _p, C.c_longlong, _i, _p, _p, _p]),
    "scda_conv3x3_dgrad_bf16_nhwc": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _i, _p, _p]),
    "scda_maxpool2x2_nhwc_bf16": (_i, [_i, _i, _i, _i, _p, _p, _p]),
    "scda_maxpool2x2_bwd_nhwc_bf16": (_i, [_i, _i, _i, _i, _p, _p, _p, _i, _p]),
    "scda_nchw_f32_to_nhwc_bf16": (_i, [_i, _i, _i, _i, _i, _p, _p, _p]),
    "scda_nhwc_bf16_to_nchw_f32": (_i, [_i, _i, _i, _i, _p, _p, _p]),
    "scda_reduce_slabs_f32": (_i, [_p, C.c_longlong, _i, _p, C.c_longlong, _i, _p]),
    "scda_colsum_bf16": (_i, [C.c_longlong, _i, _p, C.c_longlong, _p, _p]),
    "scda_colsum_f32": (_i, [C.c_longlong, _i, _p, _p, _p]),
    "scda_conv3x3_bf16_nhwc": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p, _p]),
    "scda_gemm_bf16_nn": (_i, [_i, _i, _i, _p, C.c_longlong, _p, C.c_longlong, _p, _p, C.c_longlong, _i, _p, _p, _p]),
    "scda_linear_wgrad_bf16": (_i, [_i, _i, _i, _p, C.c_longlong, _p, C.c_longlong, _p, C.c_longlong, _i, _p]),
    "scda_conv3x3_wgrad_bf16_nhwc": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _i, _p]),
    "scda_instnorm_workspace_bytes": (_z, [_i, _i, _i]),
    "scda_instnorm_act_fwd_nhwc_f32": (_i, [_i, _i, _i, _p, _p, _p, _p, _f, _i, _f, _p, _z, _p]),
    "scda_instnorm_act_bwd_nhwc_f32": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _i, _f, _p, _z, _p]),
    "scda_upsample_bilinear2x_nhwc_f32": (_i, [_i, _i, _i, _i, _p, _p, _p]),
    "scda_instnorm_act_fwd_nhwc": (_i, [_i, _i, _i, _p, _p, _i, _p, _p, _f, _i, _f, _p, _z, _p]),
    "scda_instnorm_act_bwd_nhwc": (_i, [_i, _i, _i, _p, _p, _i, _p, _p, _p, _i, _i, _f, _p, _z, _p]),
    "scda_upsample_bilinear2x_nhwc": (_i, [_i, _i, _i, _i, _p, _p, _i, _p]),
    "scda_upsample_bilinear2x_bwd_nhwc": (_i, [_i, _i, _i, _i, _p, _i, _p, _p]),
    "scda_upsample_bilinear2x_bwd_nhwc_f32": (_i, [_i, _i, _i, _i, _p, _p, _p]),
    "scda_roi_pool_nhwc_bf16_fwd": (_i, [_p, _f, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "scda_roi_pool_nhwc_bf16_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "scda_roi_pool_nhwc_f32_fwd": (_i, [_p, _f, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "scda_roi_pool_nhwc_f32_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "scda_split3_f32_bf16": (_i, [C.c_longlong, _i, _p, C.c_longlong, _p, _p]),
    "scda_split_weights_f32_bf16": (_i, [C.c_longlong, _i, _p, C.c_longlong, _p, _p, _p]),
    "scda_conv3x3_wgrad_bf16_nhwc_ld": (_i, [_i, _i, _i, _i, _i, _p, C.c_longlong, _p, C.c_longlong, _p, _i, _p]),
    "scda_maxpool2x2_nhwc_f32": (_i, [_i, _i, _i, _i, _p, _p, _p]),
    "scda_maxpool2x2_bwd_nhwc_f32": (_i, [_i, _i, _i, _i, _p, _p, _p, _i, _p]),
    "scda_nchw_f32_to_nhwc_f32": (_i, [_i, _i, _i, _i, _i, _p, _p, _p]),
    "scda_colsum_f32_ld": (_i, [C.c_longlong, _i, _p, C.c_longlong, _p, _p]),
    "scda_kmeans_workspace_bytes": (_z, [_i, _i]),
    "scda_kmeans_regions": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _f, _p, _i, _p, _p, _p, _p, _p, _z, _p]),
    "scda_smooth_l1_sigma_sum_fwd": (_i, [C.c_longlong, _p, _p, _p, _f, _p, _p]),
    "scda_smooth_l1_sigma_sum_bwd": (_i, [C.c_longlong, _p, _p, _p, _f, _p, _p, _p]),
    "scda_bce_sigmoid_rows_fwd": (_i, [_i, _i, _p, _p, _i, _p, _p]),
    "scda_bce_sigmoid_rows_bwd": (_i, [_i, _i, _p, _p, _i, _p, _p, _p]),
    "scda_conv_s2_weights": (_i, [_i, _i, _p, _p, _p]),
    "scda_conv3x3_s2_bf16_nhwc": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _f, _p]),
    "scda_conv3x3_s2_dgrad_bf16_nhwc": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _i, _p, _f, _p]),
    "scda_conv3x3_s2_wgrad_bf16_nhwc": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _i, _p]),
    "scda_conv_s2_wgrad_gather": (_i, [_i, _i, _p, _i, _p, _i, _p]),
    "scda_disc_l1_fwd": (_i, [_i, _i, _i, _p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, _p, _p, _f, _p, _i, _p]),
    "scda_disc_l1_workspace_bytes": (_z, [_i, _i, _i]),
    "scda_disc_l1_bwd": (_i, [_i, _i, _i, _p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, _p, _p, _i, _p, _p, _p,
                              _i, _p, _z, _p]),
    "scda_head_dot_fwd": (_i, [C.c_longlong, _i, _p, _i, _p, _p, _p, _p]),
    "scda_head_dot_bwd": (_i, [C.c_longlong, _i, _p, _i, _p, _p, _f, _p, _p, _p, _p]),
    "scda_bn_workspace_bytes": (_z, [C.c_longlong, _i]),
    "scda_bn_lrelu_fwd": (_i, [C.c_longlong, _i, _p, _p, _p, _f, _f, _f, _p, _p, _p, _p, _p, _i, _p, _z, _p]),
    "scda_bn_lrelu_bwd": (_i, [C.c_longlong, _i, _p, _p, _i, _p, _p, _p, _p, _f, _p, _i, _p, _p, _i, _p, _z, _p]),
    "scda_avgpool_fwd": (_i, [_i, _i, _i, _p, _p, _p]),
    "scda_avgpool_bwd": (_i, [_i, _i, _i, _p, _p, _i, _p]),
    "scda_image_prepare": (_i, [_p, _i, _i, _p, _i, _i, _i, _i, _p, _p, _p]),
    "scda_softmax_ce_workspace_bytes": (_z, [C.c_longlong]),
    "scda_softmax_ce_acc_fwd": (_i, [C.c_longlong, _i, _p, C.c_longlong, _p, C.c_longlong, _p, _p, _z, _p]),
    "scda_softmax_ce_bwd": (_i, [C.c_longlong, _i, _p, C.c_longlong, _p, C.c_longlong, _p, _p, _p, C.c_longlong, _p]),
    "scda_rpn_fg_scores": (_i, [_i, _i, _i, _i, _p, _p, _p]),
    "scda_conv1x1_tanh_workspace_bytes": (_z, [C.c_longlong, _i, _i]),
    "scda_conv1x1_tanh_fwd": (_i, [C.c_longlong, _i, _i, _p, _p, _p, _p, _p]),
    "scda_conv1x1_tanh_bwd": (_i, [C.c_longlong, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _z, _p]),
    "scda_rpn_decode_pack": (_i, [_i, _p, _p, _p, _p, C.c_double, C.c_double, C.c_double, _p, _p, _p]),
    "scda_rpn_proposal_rows_workspace_bytes": (_z, [_i, _i]),
    "scda_rpn_proposal_rows": (_i, [_i, _i, _p, _p, _p, C.c_double, C.c_double, C.c_double, _p, _p, _p, _z, _p]),
    "scda_stream_capture_id": (C.c_ulonglong, [_p]),
    "scda_timestamp": (_i, [_p, _i, _p]),
    "scda_transpose_bf16": (_i, [_i, _i, _p, C.c_longlong, _p, C.c_longlong, _p]),
    "scda_proposal_targets": (_i, [_i, _i, _p, _p, _i, _p, _i, _f, _f, _f, _f, _f, _i, _i, _i, _i, _p, _p, _f,
                                   _p, _p, _p, _p, _p, _p, _p, _p]),
    "scda_anchor_targets": (_i, [_i, _i, _i, _p, _p, _i, _p, _f, _f, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "scda_crop_regions": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "scda_conv3x3_first_nchw": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p]),
    "scda_conv3x3_set_plan": (_i, [_i, _i, _i]),
    "scda_conv3x3_set_pair": (_i, [_i]),
    "scda_conv3x3_wgrad_set_form": (_i, [_i]),
    "scda_adam_step": (_i, [_p, _p, _p, _p, _p, C.c_longlong, _i, _f, _f, _f, _f, _f, _f, _p, _p]),
}

_LIB = None


class ScdaLibraryError(RuntimeError):
    pass


def load(path: str | None = None) -> C.CDLL:
    """dlopen the library and attach prototypes.  Raises if it is not built."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ScdaLibraryError(
            "%s not found: build it with `python -m scda_b200.build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for these operators." % p)
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ScdaLibraryError("symbol %s missing from %s" % (name, p)) from e
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _LIB = lib
    return lib


# kernels launched per successful entry-point call (memsets not counted); bench.py reports
# the total inside its timed region as `gpu_launches`
KERNELS_PER_CALL = {
    "scda_nms": 2, "scda_nms_dyn": 2, "scda_instnorm_act_fwd_nhwc_f32": 3, "scda_instnorm_act_bwd_nhwc_f32": 3,
    "scda_instnorm_act_fwd_nhwc": 3, "scda_instnorm_act_bwd_nhwc": 3, "scda_conv1x1_tanh_bwd": 2,
    "scda_softmax_ce_acc_fwd": 2, "scda_disc_l1_bwd": 3, "SoftmaxFocalLossForwardLaucher": 1,
    "SoftmaxFocalLossBackwardLaucher": 1,
}
LAUNCHES = 0


def check(status: int, what: str) -> None:
    """Status convention of include/scda_b200.h: 1 ok, 0 bad arguments, <0 = -cudaError_t."""
    global LAUNCHES
    if status == 1:
        LAUNCHES += KERNELS_PER_CALL.get(what, 1)
        return
    if status == 0:
        raise ValueError("%s: arguments rejected by libscda_b200" % what)
    raise ScdaLibraryError("%s: CUDA error %d" % (what, -status))


def stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ScdaLibraryError(
                "scda_b200 operators run on CUDA tensors only (got a %s tensor); "
                "there is no CPU path" % t.device.type)
