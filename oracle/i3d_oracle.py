"""TEST INFRASTRUCTURE ONLY — CPU oracle for the clip classifier forward.

A plain PyTorch fp32 restatement of the reference network's eval-mode forward,
operating directly on a reference-schema state_dict.  It is
pinned against the UNMODIFIED reference (oracle/ref_loader.py) by
tests/golden/make_golden.py, whose outputs are committed under tests/golden/ and
re-checked by tests/test_oracle.py.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline/reference legs may import this file; the product path
never does.

Follows, op for op:
  ResNet.forward            altfreezing/slowfast/models/video_model_builder.py:561-578
  ResNetBasicStem.forward   altfreezing/slowfast/models/stem_helper.py:173-178
  ResStage/ResBlock.forward altfreezing/slowfast/models/resnet_helper.py:616-647,438-444
  BottleneckTransform.fwd   altfreezing/slowfast/models/resnet_helper.py:311-326
  ResNetBasicHead.forward   altfreezing/slowfast/models/head_helper.py:74-95
  I3D8x8.forward            altfreezing/model/classifier/i3d_ori.py:92-104
"""
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

EPS = 1e-5                       # stem_helper.py:23, resnet_helper.py:212
DEPTH = (3, 4, 6, 3)             # video_model_builder.py:18  (_MODEL_STAGE_DEPTH[50])
STRIDE = (1, 2, 2, 2)            # slowfast/config/defaults.py:164


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, EPS)


def _conv(sd, p, x, stride, pad):
    return F.conv3d(x, sd[p + ".weight"], None, stride, pad)


def stem(sd, x):
    p = "resnet.s1.pathway0_stem"
    x = F.relu(_bn(sd, p + ".bn", _conv(sd, p + ".conv", x, (1, 2, 2), (2, 3, 3))))
    return F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))


def block(sd, p, x, stride):
    kt = sd[p + ".branch2.a.weight"].shape[2]
    y = F.relu(_bn(sd, p + ".branch2.a_bn", _conv(sd, p + ".branch2.a", x, 1, (kt // 2, 0, 0))))
    y = F.relu(_bn(sd, p + ".branch2.b_bn", _conv(sd, p + ".branch2.b", y, (1, stride, stride), (0, 1, 1))))
    y = _bn(sd, p + ".branch2.c_bn", _conv(sd, p + ".branch2.c", y, 1, 0))
    if (p + ".branch1.weight") in sd:
        x = _bn(sd, p + ".branch1_bn", _conv(sd, p + ".branch1", x, (1, stride, stride), 0))
    return F.relu(x + y)


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, return_stages: bool = False):
    """x: float32 [B,3,T,H,W] normalised clip -> logits [B,1] (no activation).
    With return_stages also returns [s1..s5 outputs, pooled features [B,2048]]."""
    sd = {k: v.float() for k, v in sd.items()}
    stages: List[torch.Tensor] = []
    with torch.no_grad():
        x = stem(sd, x.float())
        stages.append(x)
        for si in range(4):
            if si == 1:
                x = F.max_pool3d(x, (2, 1, 1), (2, 1, 1))        # pathway0_pool
            for bi in range(DEPTH[si]):
                x = block(sd, "resnet.s%d.pathway0_res%d" % (si + 2, bi), x, STRIDE[si] if bi == 0 else 1)
            stages.append(x)
        t, h, w = x.shape[2:]
        # AvgPool3d([T/2, S/32, S/32], stride=1) on a [16,7,7] map == global mean
        pooled = F.avg_pool3d(x, (t, h, w), 1)
        feat = pooled.permute(0, 2, 3, 4, 1)
        logits = F.linear(feat, sd["resnet.head.projection.weight"], sd["resnet.head.projection.bias"])
        logits = logits.view(logits.shape[0], -1)
    if return_stages:
        stages.append(feat.reshape(feat.shape[0], -1))
        return logits, stages
    return logits


def conv_bn_act(x, w, gamma, beta, mean, var, stride, pad, relu, residual=None):
    """Single conv + eval BN (+residual) (+ReLU); the per-layer oracle the CUDA
    conv kernels are checked against on small shapes."""
    y = F.conv3d(x.float(), w.float(), None, stride, pad)
    y = F.batch_norm(y, mean, var, gamma, beta, False, 0.0, EPS)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y
