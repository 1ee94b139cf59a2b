"""TEST INFRASTRUCTURE ONLY — CPU oracle for the FTCN-TT plugin's forward.

A plain PyTorch fp32 restatement of the eval-mode forward of the reference's second classifier plugin,
`i3d_temporal_var_fix_dropout_tt_cfg` with setting/ftcn_tt.yaml, operating directly on a reference-schema
state_dict (275 keys).  Pinned against the UNMODIFIED reference by tests/golden/make_golden_ftcn.py (which needs
an environment shim: the plugin copies every name of nn.Conv3d's signature off the module, and torch >= 1.9 added
`device`/`dtype`, so the reference only constructs under its pinned torch 1.8 or with those two attributes present).
Only tests/ may import this file; the product path never does.

Follows, op for op:
  temporal_only_conv            altfreezing/model/classifier/i3d_temporal_var_fix_dropout_tt_cfg.py:207-289
                                (every spatial kernel -> 1, every spatial stride -> MaxPool3d((1,2,2)) behind the BN)
  I3D8x8.__init__ / forward     :295-352 (s5 = Identity for stop_point 5, head = TransformerHead(14, 16, 1024))
  ResNetBasicStem.forward       altfreezing/slowfast/models/stem_helper.py:173-178   (conv, bn[+pool], relu, pool)
  BottleneckTransform.forward   altfreezing/slowfast/models/resnet_helper.py:311-326 (a, a_bn, relu, b, b_bn[+pool], relu, c, c_bn)
  ResBlock.forward              resnet_helper.py:438-444
  TransformerHead.forward       i3d_temporal_var_fix_dropout_tt_cfg.py:183-196
  TimeTransformer.forward       altfreezing/model/classifier/time_transformer.py:259-273 (Transformer :75-88,
                                Attention :29-73, FeedForward :17-28, PreNorm/Residual :8-16)
"""
from typing import Dict, List

import torch
import torch.nn.functional as F

EPS = 1e-5
DEPTH = (3, 4, 6)                # s2..s4 (s5 is nn.Identity)
STRIDED = (False, True, True)    # stages whose first block lost a spatial stride to a max-pool
HEADS, DIM_HEAD = 16, 64
TT = "resnet.head.time_T"


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, EPS)


def _conv(sd, p, x, pad):
    return F.conv3d(x, sd[p + ".weight"], None, 1, pad)


def _pool2(x):
    return F.max_pool3d(x, (1, 2, 2))


def stem(sd, x):
    p = "resnet.s1.pathway0_stem"
    x = F.relu(_pool2(_bn(sd, p + ".bn.0", _conv(sd, p + ".conv", x, (2, 0, 0)))))
    return F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))


def block(sd, p, x, pooled):
    kt = sd[p + ".branch2.a.weight"].shape[2]
    y = F.relu(_bn(sd, p + ".branch2.a_bn", _conv(sd, p + ".branch2.a", x, (kt // 2, 0, 0))))
    y = _conv(sd, p + ".branch2.b", y, 0)
    y = F.relu(_pool2(_bn(sd, p + ".branch2.b_bn.0", y)) if pooled else _bn(sd, p + ".branch2.b_bn", y))
    y = _bn(sd, p + ".branch2.c_bn", _conv(sd, p + ".branch2.c", y, 0))
    if (p + ".branch1.weight") in sd:
        s = _conv(sd, p + ".branch1", x, 0)
        x = _pool2(_bn(sd, p + ".branch1_bn.0", s)) if pooled else _bn(sd, p + ".branch1_bn", s)
    return F.relu(x + y)


def transformer_head(sd, tokens, return_cls=False):
    """tokens [B, n, D] -> logits [B, 1]."""
    b, n, d = tokens.shape
    x = torch.cat((sd[TT + ".cls_token"].expand(b, -1, -1), tokens), dim=1) + sd[TT + ".pos_embedding"][:, : n + 1]
    i = 0
    while (TT + ".transformer.layers.%d.0.fn.norm.weight" % i) in sd:
        q = TT + ".transformer.layers.%d" % i
        h = F.layer_norm(x, (d,), sd[q + ".0.fn.norm.weight"], sd[q + ".0.fn.norm.bias"], EPS)
        qkv = F.linear(h, sd[q + ".0.fn.fn.to_qkv.weight"]).chunk(3, dim=-1)
        qh, kh, vh = [t.reshape(b, n + 1, HEADS, DIM_HEAD).permute(0, 2, 1, 3) for t in qkv]
        att = (torch.einsum("bhid,bhjd->bhij", qh, kh) * DIM_HEAD ** -0.5).softmax(dim=-1)
        o = torch.einsum("bhij,bhjd->bhid", att, vh).permute(0, 2, 1, 3).reshape(b, n + 1, HEADS * DIM_HEAD)
        x = F.linear(o, sd[q + ".0.fn.fn.to_out.0.weight"], sd[q + ".0.fn.fn.to_out.0.bias"]) + x
        h = F.layer_norm(x, (d,), sd[q + ".1.fn.norm.weight"], sd[q + ".1.fn.norm.bias"], EPS)
        h = F.gelu(F.linear(h, sd[q + ".1.fn.fn.net.0.weight"], sd[q + ".1.fn.fn.net.0.bias"]))
        x = F.linear(h, sd[q + ".1.fn.fn.net.3.weight"], sd[q + ".1.fn.fn.net.3.bias"]) + x
        i += 1
    cls = F.layer_norm(x[:, 0], (d,), sd[TT + ".mlp_head.0.weight"], sd[TT + ".mlp_head.0.bias"], EPS)
    logits = F.linear(cls, sd[TT + ".mlp_head.1.weight"], sd[TT + ".mlp_head.1.bias"])
    return (logits, cls) if return_cls else logits


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, return_stages: bool = False):
    """x: float32 [B,3,T,H,W] normalised clip -> logits [B,1] (no activation).  With return_stages also returns
    [s1..s4 outputs, tokens [B,16,1024], normalised cls vector [B,1024]]."""
    sd = {k: v.float() for k, v in sd.items()}
    stages: List[torch.Tensor] = []
    with torch.no_grad():
        x = stem(sd, x.float())
        stages.append(x)
        for si in range(3):
            if si == 1:
                x = F.max_pool3d(x, (2, 1, 1), (2, 1, 1))        # pathway0_pool
            for bi in range(DEPTH[si]):
                x = block(sd, "resnet.s%d.pathway0_res%d" % (si + 2, bi), x, STRIDED[si] and bi == 0)
            stages.append(x)
        b, c, t, h, w = x.shape
        tokens = F.avg_pool3d(x, (1, h, w)).reshape(-1, c, t).permute(0, 2, 1)
        logits, cls = transformer_head(sd, tokens, return_cls=True)
        logits = logits.reshape(b, -1)
    if return_stages:
        stages += [tokens.contiguous(), cls]
        return logits, stages
    return logits
