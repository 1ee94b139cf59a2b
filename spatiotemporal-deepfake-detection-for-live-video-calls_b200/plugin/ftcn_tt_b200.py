"""Reference-side plugin for the FTCN-TT classifier: copy (or symlink) this file to
`altfreezing/model/classifier/ftcn_tt_b200.py` and set `classifier_type: ftcn_tt_b200` in a copy of
setting/ftcn_tt.yaml.  `PluginLoader.get_classifier("ftcn_tt_b200")` then returns this
`Classifier` (altfreezing/utils/plugin_loader.py:27-30,42-52).

It builds the reference's own FTCN-TT `I3D8x8` (model/classifier/i3d_temporal_var_fix_dropout_tt_cfg.py:295-333)
as `self.network`, so `load()`, `state_dict()` and checkpoint keys are exactly the reference's, and swaps the
B200 engine (variant "ftcn_tt") into the `_warped_network` slot that `ModelBase.forward` calls
(altfreezing/model/_base.py:22-26).  Note that the reference module itself only constructs under torch 1.8 (it
copies every name of nn.Conv3d's signature off the module, :198,238); on newer torch give nn.Conv3d the two
attributes it looks for before importing it: `nn.Conv3d.device = None; nn.Conv3d.dtype = None`.
"""
import os
import sys

from .i3d_temporal_var_fix_dropout_tt_cfg import I3D8x8      # the reference network definition
from ._classifier_base import ClassifierBase

_REPO = os.environ.get("AFB200_ROOT")
if _REPO and _REPO not in sys.path:
    sys.path.insert(0, _REPO)
import afb200                                        # noqa: E402,F401
from afb200.classifier import B200Engine             # noqa: E402


class Classifier(ClassifierBase):
    @property
    def module_to_build(self):
        return I3D8x8

    def __init__(self):
        super().__init__()
        from config import config as cfg
        engine = B200Engine(self.network, precision=os.environ.get("AFB200_PRECISION", "bf16"),
                            max_batch=int(os.environ.get("AFB200_MAX_BATCH", "32")),
                            clip_t=cfg.clip_size, clip_s=cfg.imsize, variant="ftcn_tt")
        object.__setattr__(self, "_warped_network", engine)

    def load(self, *args, **kwargs):
        out = super().load(*args, **kwargs)
        self._warped_network.refold()
        return out
