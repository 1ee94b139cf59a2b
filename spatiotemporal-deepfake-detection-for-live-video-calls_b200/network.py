"""Checkpoint container with the reference's state_dict schema.

`I3D8x8Params` registers exactly the 320 tensors of the reference `I3D8x8` network
(altfreezing/model/classifier/i3d_ori.py:73-90 wrapping slowfast ResNet;
SURVEY.md App. C) under the same names, so reference checkpoints load with
`load_state_dict` unchanged.  It holds parameters only: its forward runs the CUDA
engine, never PyTorch ops.  `resnet.head.projection` is a real nn.Linear so that
altfreezing/feature.py:106-114 (hook on the last nn.Linear) keeps working.
"""
from typing import Dict

import torch
from torch import nn

from . import arch


class _Holder(nn.Module):
    """An empty module used to reproduce the reference's attribute paths."""


def _conv_holder(spec: arch.ConvSpec) -> nn.Module:
    m = _Holder()
    m.weight = nn.Parameter(torch.zeros((spec.cout, spec.cin) + tuple(spec.kernel)), requires_grad=False)
    return m


def _bn_holder(c: int) -> nn.Module:
    m = _Holder()
    m.weight = nn.Parameter(torch.ones(c), requires_grad=False)
    m.bias = nn.Parameter(torch.zeros(c), requires_grad=False)
    m.register_buffer("running_mean", torch.zeros(c))
    m.register_buffer("running_var", torch.ones(c))
    m.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
    return m


def _attach(root: nn.Module, dotted: str, leaf: nn.Module):
    parts = dotted.split(".")
    cur = root
    for p in parts[:-1]:
        if not hasattr(cur, p):
            setattr(cur, p, _Holder())
        cur = getattr(cur, p)
    setattr(cur, parts[-1], leaf)


class I3D8x8Params(nn.Module):
    def __init__(self):
        super().__init__()
        for spec in arch.all_conv_specs():
            _attach(self, spec.name, _conv_holder(spec))
            _attach(self, spec.bn, _bn_holder(spec.cout))
        _attach(self, "resnet.head.projection", nn.Linear(arch.FEATURE_DIM, 1, bias=True))

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("I3D8x8Params holds weights only; run it through afb200.B200Engine")


class _ParamHolder(_Holder):
    """A leaf that owns one tensor per given attribute name (LayerNorm / bias-free Linear stand-ins)."""

    def __init__(self, **shapes):
        super().__init__()
        for name, shape in shapes.items():
            setattr(self, name, nn.Parameter(torch.zeros(shape), requires_grad=False))


class FTCNTTParams(nn.Module):
    """The 275 tensors of the reference FTCN-TT network (model/classifier/i3d_temporal_var_fix_dropout_tt_cfg.py:
    295-333 with setting/ftcn_tt.yaml) under the reference's names — BatchNorms that temporal_only_conv wrapped into
    nn.Sequential(bn, MaxPool3d) carry the extra ".0" — so its checkpoints load unchanged.  The head's final
    nn.Linear (mlp_head.1) is a real nn.Linear for altfreezing/feature.py:106-114's hook."""

    def __init__(self):
        super().__init__()
        for spec in arch.all_conv_specs("ftcn_tt"):
            _attach(self, spec.name, _conv_holder(spec))
            _attach(self, spec.bn, _bn_holder(spec.cout))
        groups: Dict[str, Dict[str, tuple]] = {}
        for name, shape in arch.tt_param_shapes().items():
            owner, leaf = name.rsplit(".", 1)
            groups.setdefault(owner, {})[leaf] = shape
        for owner, shapes in groups.items():
            if owner.endswith("mlp_head.1"):
                _attach(self, owner, nn.Linear(arch.TT_DIM, 1, bias=True))
            elif owner == arch.TT_PREFIX:          # pos_embedding / cls_token live on time_T itself, next to children
                holder = self
                for part in owner.split("."):
                    if not hasattr(holder, part):
                        setattr(holder, part, _Holder())
                    holder = getattr(holder, part)
                for leaf, shape in shapes.items():
                    setattr(holder, leaf, nn.Parameter(torch.zeros(shape), requires_grad=False))
            else:
                _attach(self, owner, _ParamHolder(**shapes))

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("FTCNTTParams holds weights only; run it through afb200.B200Engine")


def params_for(variant: str) -> nn.Module:
    return FTCNTTParams() if variant == "ftcn_tt" else I3D8x8Params()


def reference_key_set(variant: str = "i3d") -> Dict[str, tuple]:
    return {k: tuple(v.shape) for k, v in params_for(variant).state_dict().items()}
