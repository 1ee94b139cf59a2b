"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

Loads the UNMODIFIED reference implementation from /root/reference so that the
restatements in oracle/ can be pinned against it and golden vectors generated
(tests/golden/make_golden.py).  /root/reference exists only in the build
container, never on the GPU box, so nothing under `-m gpu`, smoke() or bench.py
may call into this module.

Three pure-Python dependencies of the reference are absent from this image
(fvcore, simplejson, termcolor); they are stubbed in sys.modules with the
minimum surface the reference touches:
  * fvcore.common.config.CfgNode      <- altfreezing/slowfast/config/defaults.py:6,23-27
  * fvcore.common.registry.Registry   <- altfreezing/slowfast/models/build.py:9
  * fvcore.nn.weight_init.c2_msra_fill<- altfreezing/slowfast/utils/weight_init_helper.py:7
  * fvcore.common.file_io.PathManager <- altfreezing/slowfast/utils/logging.py:13
  * simplejson, termcolor             <- slowfast/utils/logging.py:12, altfreezing/utils/logger.py:32
  * timm.models.layers.trunc_normal_  <- model/classifier/time_transformer.py:217 (FTCN-TT plugin only)
"""
import copy
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AFB200_REFERENCE_ROOT", "/root/reference")
ALTFREEZING_DIR = os.path.join(REFERENCE_ROOT, "altfreezing")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(ALTFREEZING_DIR, "slowfast"))


class _CfgNode(dict):
    """yacs-like attribute dict: nested dicts become nodes; clone(); merge."""

    def __init__(self, init=None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _CfgNode):
            v = type(self)(v) if type(self) is not _CfgNode else _CfgNode(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_other_cfg(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].merge_from_other_cfg(v)
            else:
                self[k] = copy.deepcopy(v)


class _Registry:
    def __init__(self, name):
        self._name, self._map = name, {}

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._map[o.__name__] = o
                return o
            return deco
        self._map[obj.__name__] = obj
        return obj

    def get(self, name):
        return self._map[name]


def _install_stubs():
    import torch.nn as nn

    def c2_msra_fill(m):
        nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        if getattr(m, "bias", None) is not None:
            nn.init.constant_(m.bias, 0)

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("fvcore")
    mod("fvcore.common")
    mod("fvcore.common.config", CfgNode=_CfgNode)
    mod("fvcore.common.registry", Registry=_Registry)
    mod("fvcore.common.file_io", PathManager=object())
    mod("fvcore.nn")
    mod("fvcore.nn.weight_init", c2_msra_fill=c2_msra_fill)
    mod("simplejson")
    mod("termcolor", colored=lambda s, *a, **k: s)
    # timm is only used for its truncated-normal initialiser (model/classifier/time_transformer.py:217,253)
    mod("timm")
    mod("timm.models")
    mod("timm.models.layers", trunc_normal_=nn.init.trunc_normal_)


_CLASSIFIER_CLS = None
_CLASSIFIER_SETTING = None


def reference_classifier(setting: str = "i3d_ori.yaml"):
    """Build `PluginLoader.get_classifier(cfg.classifier_type)()` exactly as
    altfreezing/demo.py:398-404 does.  `setting` is the yaml under altfreezing/setting/:
    "i3d_ori.yaml" (AltFreezing I3D) or "ftcn_tt.yaml" (the FTCN-TT plugin).  cfg is a
    process-wide singleton and freeze() is irreversible, so one process can build one setting only."""
    global _CLASSIFIER_CLS, _CLASSIFIER_SETTING
    if not reference_available():
        raise RuntimeError("reference tree not present: " + REFERENCE_ROOT)
    if _CLASSIFIER_CLS is not None and _CLASSIFIER_SETTING != setting:
        raise RuntimeError("the reference config is frozen on %s in this process; build %s in another process"
                           % (_CLASSIFIER_SETTING, setting))
    if _CLASSIFIER_CLS is None:
        _install_stubs()
        if ALTFREEZING_DIR not in sys.path:
            sys.path.insert(0, ALTFREEZING_DIR)
        from config import config as cfg
        cfg.init_with_yaml()
        cfg.update_with_yaml(setting)
        cfg.freeze()
        from utils.plugin_loader import PluginLoader
        _CLASSIFIER_CLS = PluginLoader.get_classifier(cfg.classifier_type)
        _CLASSIFIER_SETTING = setting
    return _CLASSIFIER_CLS().eval()


def reference_crop_align(size=224):
    """The reference FasterCropAlignXRay (numpy + cv2.warpAffine),
    altfreezing/test_tools/faster_crop_align_xray.py:11-88."""
    if ALTFREEZING_DIR not in sys.path:
        sys.path.insert(0, ALTFREEZING_DIR)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from test_tools.faster_crop_align_xray import FasterCropAlignXRay
    return FasterCropAlignXRay(size)


def reference_get_crop_box():
    if ALTFREEZING_DIR not in sys.path:
        sys.path.insert(0, ALTFREEZING_DIR)
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_ref_tt_utils", os.path.join(ALTFREEZING_DIR, "test_tools", "utils.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.get_crop_box
