"""afb200 — B200-native AltFreezing clip-classification hot path.

Host side (Python) of the C-ABI library in csrc/ (include/afb200.h).  Importing the
package does not load the CUDA library; the first engine/crop call does, and raises if the
built `libafb200.so` is missing — there is no CPU fallback.
"""
from . import arch, features, live, parallel, report, synthetic  # noqa: F401
from ._lib import Afb200Error, LIB_PATH, lib  # noqa: F401
from .classifier import B200Engine, Classifier, RGBBackboneB200  # noqa: F401
from .crop import CropAlignB200, clip_geometry, clip_geometry_batch, estimate_clip_transform, get_crop_box, get_crop_boxes  # noqa: F401
from .engine import Engine, conv_bc_fused_ndhwc, conv_ndhwc, conv_shortcut_ndhwc, mean_std_255, set_global_option, stem_pool_ndhwc4  # noqa: F401
from .network import I3D8x8Params  # noqa: F401
from .service import ClassifierSvc, CropAlignSvc  # noqa: F401
from .weights import FoldedWeights, fold_conv_bn, strip_checkpoint  # noqa: F401

__all__ = ["arch", "synthetic", "lib", "Engine", "B200Engine", "Classifier", "RGBBackboneB200", "CropAlignB200", "ClassifierSvc",
           "CropAlignSvc", "I3D8x8Params", "FoldedWeights", "conv_ndhwc", "get_crop_box", "clip_geometry",
           "estimate_clip_transform", "mean_std_255", "Afb200Error"]
