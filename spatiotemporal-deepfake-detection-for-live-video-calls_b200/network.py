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


def reference_key_set() -> Dict[str, tuple]:
    return {k: tuple(v.shape) for k, v in I3D8x8Params().state_dict().items()}
