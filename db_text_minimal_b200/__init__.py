"""db_text_minimal_b200 -- B200 (sm_100a) native hot path of DB_text_minimal.

Drop-in mirrors of the reference's hot-path classes (same names, arguments and return values):
    DBTextModel                      <- src/models.py
    DBLoss, OHEMBalanceCrossEntropyLoss, DiceLoss, L1Loss   <- src/losses.py
    SegDetectorRepresenter           <- src/postprocess.py
Everything below them runs in hand-written CUDA kernels reached through the C ABI of
libdbb200.so (include/dbb200.h).  There is no CPU / PyTorch-eager fallback: calls on non-CUDA
tensors, or with the library missing, raise.
"""
from ._lib import DbbError, lib  # noqa: F401

__all__ = ["DbbError", "lib"]


def __getattr__(name):
    # lazy: importing the package must work on a CPU-only box (the driver's build check)
    if name in ("DBLoss", "OHEMBalanceCrossEntropyLoss", "DiceLoss", "L1Loss"):
        from . import losses
        return getattr(losses, name)
    if name in ("DBTextModel", "backbone_dict", "segmentation_body_dict", "segmentation_head_dict"):
        from . import models
        return getattr(models, name)
    if name == "SegDetectorRepresenter":
        from . import postprocess
        return postprocess.SegDetectorRepresenter
    raise AttributeError(name)
