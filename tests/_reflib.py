"""Binding of oracle/_ref/libscda_ref.so — the reference's own .cu files compiled
unmodified (oracle/build.py).  Its launcher symbols are the Section A names of
include/scda_b200.h, so the prototypes are shared with scda_b200._lib."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libscda_ref.so")

SECTION_A = [
    "ROIPoolForwardLaucher", "ROIPoolBackwardLaucher", "ROIAlignForwardLaucher",
    "ROIAlignBackwardLaucher", "_nms", "IOUOverlap", "SigmoidFocalLossForwardLaucher",
    "SigmoidFocalLossBackwardLaucher", "SoftmaxFocalLossForwardLaucher",
    "SoftmaxFocalLossBackwardLaucher",
]


def available():
    return os.path.exists(REF_SO)


def load():
    from scda_b200._lib import SIGNATURES
    lib = C.CDLL(REF_SO)
    for name in SECTION_A:
        res, args = SIGNATURES[name]
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
