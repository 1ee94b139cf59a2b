"""Reference checkpoint -> folded kernel weights.

Takes a state_dict in the reference schema (320 keys) and folds every
eval-mode BatchNorm3d into its Conv3d:
    s = gamma / sqrt(running_var + eps);  W' = W * s[:,None,None,None,None];  b' = beta - running_mean * s
(eps = 1e-5: altfreezing/slowfast/models/stem_helper.py:23, resnet_helper.py:212).
The result is what `af_create` uploads (include/afb200.h: af_weights).
"""
import ctypes as C
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import arch
from ._lib import AfBlockDesc, AfConvDesc, AfTTHead, AfTTLayer, AfWeights


def strip_checkpoint(saved) -> Dict[str, torch.Tensor]:
    """The tolerant unwrapping of ModelBase.load (altfreezing/model/_base.py:59-73):
    accept a bare state_dict or one wrapped under state_dict / classifier_state_dict /
    model_state_dict, and strip ONE leading module. / network. / _warped_network. prefix."""
    sd = saved
    if isinstance(saved, dict):
        for k in ("state_dict", "classifier_state_dict", "model_state_dict"):
            if k in saved:
                sd = saved[k]
                break

    def strip(k):
        for p in ("module.", "network.", "_warped_network."):
            if k.startswith(p):
                return k[len(p):]
        return k
    return {strip(k): v for k, v in sd.items()}


def fold_conv_bn(sd: Dict[str, torch.Tensor], spec: arch.ConvSpec) -> Tuple[np.ndarray, np.ndarray]:
    w = sd[spec.name + ".weight"].detach().double().cpu()
    g = sd[spec.bn + ".weight"].detach().double().cpu()
    b = sd[spec.bn + ".bias"].detach().double().cpu()
    m = sd[spec.bn + ".running_mean"].detach().double().cpu()
    v = sd[spec.bn + ".running_var"].detach().double().cpu()
    s = g / torch.sqrt(v + arch.BN_EPS)
    wf = (w * s.view(-1, 1, 1, 1, 1)).float().contiguous().numpy()
    bf = (b - m * s).float().contiguous().numpy()
    assert wf.shape == (spec.cout, spec.cin) + tuple(spec.kernel), (spec.name, wf.shape)
    return wf, bf


class FoldedWeights:
    """Owns the folded host arrays and the ctypes structures pointing into them.
    variant "i3d": the AltFreezing I3D (i3d_ori plugin); "ftcn_tt": the FTCN-TT plugin (arch.py)."""

    def __init__(self, sd: Dict[str, torch.Tensor], clip_t: int = 32, clip_s: int = 224, variant: str = "i3d"):
        if variant not in arch.VARIANTS:
            raise ValueError("unknown network variant %r (one of %s)" % (variant, arch.VARIANTS))
        self.variant = variant
        self.specs: List[arch.ConvSpec] = arch.all_conv_specs(variant)
        index = {s.name: i for i, s in enumerate(self.specs)}
        self._arrays = []
        self.convs = (AfConvDesc * len(self.specs))()
        for i, sp in enumerate(self.specs):
            w, b = fold_conv_bn(sd, sp)
            self._arrays += [w, b]
            d = self.convs[i]
            d.weight = w.ctypes.data
            d.bias = b.ctypes.data
            d.cin, d.cout = sp.cin, sp.cout
            d.kt, d.kh, d.kw = sp.kernel
            d.st, d.sh, d.sw = sp.stride
            d.pt, d.ph, d.pw = sp.pad
        blocks = arch.block_specs_for(variant)
        self.blocks = (AfBlockDesc * len(blocks))()
        for i, blk in enumerate(blocks):
            bd = self.blocks[i]
            bd.branch1 = index[blk.branch1.name] if blk.branch1 is not None else -1
            bd.a, bd.b, bd.c = index[blk.a.name], index[blk.b.name], index[blk.c.name]
            bd.temporal_pool_before = 1 if (blk.stage == 3 and blk.index == 0) else 0
            bd.spatial_pool = 1 if blk.b.pool2 else 0
        self.struct = AfWeights()
        self.struct.n_convs = len(self.specs)
        self.struct.convs = C.cast(self.convs, C.POINTER(AfConvDesc))
        stem = arch.stem_spec_for(variant)
        self.struct.stem = index[stem.name]
        self.struct.stem_pool2 = 1 if stem.pool2 else 0
        self.struct.n_blocks = len(blocks)
        self.struct.blocks = C.cast(self.blocks, C.POINTER(AfBlockDesc))
        self.struct.feature_dim = arch.feature_dim_for(variant)
        self.struct.clip_t, self.struct.clip_s = clip_t, clip_s
        if variant == "ftcn_tt":
            self._build_tt_head(sd)
        else:
            self.fc_w = sd["resnet.head.projection.weight"].detach().float().cpu().contiguous().numpy().reshape(-1)
            self.fc_b = float(sd["resnet.head.projection.bias"].detach().float().cpu().reshape(-1)[0])
            assert self.fc_w.shape[0] == arch.FEATURE_DIM
            self.struct.fc_weight = self.fc_w.ctypes.data
            self.struct.fc_bias = self.fc_b

    def _build_tt_head(self, sd):
        """TransformerHead / TimeTransformer parameters (arch.tt_param_shapes) -> af_tt_head."""
        shapes = arch.tt_param_shapes()

        def arr(name):
            a = sd[name].detach().float().cpu().contiguous().numpy()
            assert tuple(a.shape) == shapes[name], (name, a.shape, shapes[name])
            a = np.ascontiguousarray(a.reshape(-1))
            self._arrays.append(a)
            return a.ctypes.data

        P = arch.TT_PREFIX
        self.tt_layers = (AfTTLayer * arch.TT_DEPTH)()
        for i in range(arch.TT_DEPTH):
            q = "%s.transformer.layers.%d" % (P, i)
            L = self.tt_layers[i]
            L.ln1_w, L.ln1_b = arr(q + ".0.fn.norm.weight"), arr(q + ".0.fn.norm.bias")
            L.qkv_w = arr(q + ".0.fn.fn.to_qkv.weight")
            L.out_w, L.out_b = arr(q + ".0.fn.fn.to_out.0.weight"), arr(q + ".0.fn.fn.to_out.0.bias")
            L.ln2_w, L.ln2_b = arr(q + ".1.fn.norm.weight"), arr(q + ".1.fn.norm.bias")
            L.fc1_w, L.fc1_b = arr(q + ".1.fn.fn.net.0.weight"), arr(q + ".1.fn.fn.net.0.bias")
            L.fc2_w, L.fc2_b = arr(q + ".1.fn.fn.net.3.weight"), arr(q + ".1.fn.fn.net.3.bias")
        h = AfTTHead()
        h.dim, h.tokens, h.heads, h.dim_head = arch.TT_DIM, arch.TT_TOKENS, arch.TT_HEADS, arch.TT_DIM_HEAD
        h.mlp_dim, h.depth = arch.TT_MLP, arch.TT_DEPTH
        h.layers = C.cast(self.tt_layers, C.POINTER(AfTTLayer))
        h.cls_token, h.pos_embedding = arr(P + ".cls_token"), arr(P + ".pos_embedding")
        h.norm_w, h.norm_b = arr(P + ".mlp_head.0.weight"), arr(P + ".mlp_head.0.bias")
        h.fc_w = arr(P + ".mlp_head.1.weight")
        h.fc_b = float(sd[P + ".mlp_head.1.bias"].detach().float().cpu().reshape(-1)[0])
        self.tt_head = h
        self.struct.tt_head = C.pointer(h)
        self.struct.fc_weight = None
        self.struct.fc_bias = 0.0
