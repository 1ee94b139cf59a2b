"""The classifier plugin behind the reference's interface.

`B200Engine` is the nn.Module that takes the place of `ModelBase._warped_network`
(altfreezing/model/_base.py:22-26): `engine(x)` -> {"final_output": logits [B,1]}.
`Classifier` mirrors what callers do with the reference plugin
(`PluginLoader.get_classifier(name)()` -> `.to(dev).eval()` / `.cuda()` -> `.load(ckpt)` ->
`clf(x)["final_output"]`; altfreezing/demo.py:403-404,323-328, TEST2.py:142-143,166-175,
test/af_realtime.py:68-69,84-89) without needing the reference tree; the file
plugin/i3d_b200.py is the variant that subclasses the reference's own ClassifierBase.
"""
import logging
from typing import Optional

import torch
from torch import nn

from .engine import Engine
from .network import params_for
from .weights import strip_checkpoint

log = logging.getLogger("afb200")


class B200Engine(nn.Module):
    """Runs `network`'s weights through libafb200.  `network` is referenced, not
    registered (the reference's ModelBase already owns it as `self.network`; registering
    it again would duplicate every key in state_dict, SURVEY.md §8b gotcha (i))."""

    def __init__(self, network: nn.Module, precision: str = "bf16", max_batch: int = 32,
                 clip_t: int = 32, clip_s: int = 224, variant: str = "i3d"):
        super().__init__()
        object.__setattr__(self, "_network", network)
        self.precision, self.max_batch, self.clip_t, self.clip_s = precision, max_batch, clip_t, clip_s
        self.variant = variant          # "i3d" (i3d_ori plugin) or "ftcn_tt" (i3d_temporal_var_fix_dropout_tt_cfg plugin)
        self._engine: Optional[Engine] = None
        self._engine_device = None
        self._weights_version = None
        self._feature_hooks = False

    def _version(self):
        return tuple((k, v._version, v.data_ptr()) for k, v in self._network.state_dict().items() if v.dim() > 0)

    def refold(self):
        """Drop the device copy of the weights; the next call re-folds from `network`
        (call after `load()` / load_state_dict)."""
        if self._engine is not None:
            self._engine.close()
        self._engine = None

    def engine_for(self, device: torch.device) -> Engine:
        ver = self._version()
        if self._engine is None or self._engine_device != device or ver != self._weights_version:
            self.refold()
            idx = device.index if device.index is not None else torch.cuda.current_device()
            self._engine = Engine(self._network.state_dict(), device=idx, max_batch=self.max_batch,
                                  precision=self.precision, clip_t=self.clip_t, clip_s=self.clip_s,
                                  variant=self.variant)
            self._engine_device, self._weights_version = device, ver
        return self._engine

    def forward(self, images, noise=None, has_mask=None, freeze_backbone=False, return_feature_maps=False):
        assert not freeze_backbone                      # as I3D8x8.forward, i3d_ori.py:100
        if not images.is_cuda:
            raise RuntimeError("afb200: the B200 engine needs a CUDA tensor (no CPU fallback); "
                               "move the classifier and its input to a B200 device")
        dev = images.device if images.device.index is not None else torch.device("cuda", torch.cuda.current_device())
        eng = self.engine_for(dev)
        proj = self._projection()
        if proj is not None and (proj._forward_hooks or proj._forward_pre_hooks):
            # feature.py:106-114 hooks the last nn.Linear: feed it the pooled features so the
            # hook observes the same input/output it would see in the reference network.
            _, feats = eng.forward(images, return_features=True)
            if self.variant == "ftcn_tt":      # mlp_head.1 sees the LayerNorm'd cls token [B, dim]
                logits = proj(feats).view(feats.shape[0], -1)
            else:
                logits = proj(feats.view(feats.shape[0], 1, 1, 1, -1)).view(feats.shape[0], -1)
        else:
            logits = eng.forward(images)
        return {"final_output": logits}

    def _projection(self):
        try:
            if self.variant == "ftcn_tt":
                return getattr(self._network.resnet.head.time_T.mlp_head, "1")
            return self._network.resnet.head.projection
        except AttributeError:
            return None


class Classifier(nn.Module):
    """Stand-alone equivalent of `PluginLoader.get_classifier("i3d_ori")` (variant "i3d") or of
    `PluginLoader.get_classifier("i3d_temporal_var_fix_dropout_tt_cfg")` (variant "ftcn_tt") on B200."""

    def __init__(self, precision: str = "bf16", max_batch: int = 32, clip_size: int = 32, imsize: int = 224,
                 variant: str = "i3d"):
        super().__init__()
        self.network = params_for(variant)
        engine = B200Engine(self.network, precision, max_batch, clip_size, imsize, variant)
        object.__setattr__(self, "_warped_network", engine)

    def forward(self, *inputs, **kwargs):
        return self._warped_network(*inputs, **kwargs)

    def parameters(self, recurse=True):            # model/_base.py:174-175
        return self.network.parameters(recurse)

    def load(self, fullpath=None, epoch=-1, pretrained=None):
        """ModelBase.load (model/_base.py:39-104): tolerant of wrappers, prefixes, unknown and
        mis-shaped keys; returns (True, epoch) or (False, -1) on ValueError/OSError."""
        if fullpath is None:
            fullpath = pretrained
        if fullpath is None:
            log.info("No existing model found")
            return False, -1
        try:
            saved = torch.load(fullpath, map_location="cpu", weights_only=False)
            self.load_state_dict_tolerant(saved)
        except (ValueError, OSError) as err:
            log.warning("Failed loading %s: %s", fullpath, err)
            return False, -1
        return True, epoch

    def load_state_dict_tolerant(self, saved):
        sd = strip_checkpoint(saved)
        own = self.network.state_dict()
        ok = {k: v for k, v in sd.items() if k in own and own[k].shape == v.shape}
        redundant = sorted(k for k in sd if k not in own)
        mismatch = sorted(k for k in sd if k in own and own[k].shape != sd[k].shape)
        missing = sorted(set(own) - set(ok) - set(mismatch))
        if redundant:
            log.warning("%s are in checkpoint, but not found in model", set(redundant))
        if missing:
            log.warning("%s are in model, but not found in checkpoint", set(missing))
        if mismatch:
            log.warning("%s have unmatching shape between checkpoint&model", set(mismatch))
        own.update(ok)
        self.network.load_state_dict(own, strict=False)
        self._warped_network.refold()
        return ok, redundant, missing, mismatch


class RGBBackboneB200(nn.Module):
    """Backbone adapter for dualrun's `AltFreezingRGBEncoder(backbone, out_dim=2048)`
    (dualrun/model/dual_rgb.py:9-44): frames `[B, T, 3, H, W]` (normalised RGB) -> per-frame features
    `[B, T/2, 2048]` from the B200 trunk.  The reference never ships such an adapter; the
    semantics chosen here make `encoder(x)` (temporal mean) equal the pooled 2048-d input of `head.projection`."""

    def __init__(self, classifier: "Classifier"):
        super().__init__()
        object.__setattr__(self, "_clf", classifier)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 5 or x.shape[2] != 3:
            raise ValueError("RGBBackboneB200 takes frames [B,T,3,H,W], got %s" % (tuple(x.shape),))
        if not x.is_cuda:
            raise RuntimeError("afb200: the B200 backbone needs a CUDA tensor (no CPU fallback)")
        eng = self._clf._warped_network
        dev = x.device if x.device.index is not None else torch.device("cuda", torch.cuda.current_device())
        logits, feats = eng.engine_for(dev).forward_frames(x.permute(0, 2, 1, 3, 4))     # strided NCTHW view
        return feats
