"""Operator package, same two names the reference exports (extensions/__init__.py:1,3)."""
from ._nms.pth_nms import pth_nms as nms
from ._roi_pooling.modules.roi_pool import _RoIPooling as RoIPool

__all__ = ["nms", "RoIPool"]
