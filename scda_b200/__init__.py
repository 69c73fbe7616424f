"""scda_b200 — B200-native (sm_100a) operator hot path of SCDA's Faster R-CNN.

The package mirrors the reference's Python surface for this path:

    scda_b200.extensions   <- extensions/   (nms, RoIPool, _roi_align, _roi_pooling,
                                             _nms, _focal_loss, _bbox_helper, _cython_bbox)
    scda_b200.functions    <- functions/    (rpn_proposal, proposal_target, anchor_target, ...)
    scda_b200.models       <- models/       (head, faster_rcnn.*)
    scda_b200.utils        <- utils/        (bbox_helper, anchor_helper, distributed_utils)

`scda_b200.compat.install()` registers those sub-packages under the reference's
top-level names so `from extensions import nms, RoIPool` resolves unchanged.

All compute goes through libscda_b200.so (include/scda_b200.h).  There is no
CPU fallback; operators raise if the library is missing or a tensor is not on
a CUDA device.
"""
__version__ = "0.1.0"
