"""Static description of the AltFreezing I3D ResNet-50 clip classifier.

This is the single source the engine, the weight folder and the checkpoint
container are generated from.  It restates what the reference builds
dynamically from its config:

  * stem        altfreezing/slowfast/models/video_model_builder.py:445-452
                (kernel [5,7,7], stride [1,2,2], pad [2,3,3], 3->64) and
                stem_helper.py:156-171 (max-pool [1,3,3]/[1,2,2]/[0,1,1])
  * stages      video_model_builder.py:454-546 with
                _MODEL_STAGE_DEPTH[50]=(3,4,6,3) (:18),
                _TEMPORAL_KERNEL_BASIS["i3d"] (:36-42),
                NUM_BLOCK_TEMP_KERNEL [[3],[4],[6],[3]] (model/classifier/i3d_ori.py:26),
                temporal-kernel expansion resnet_helper.py:530-534,
                SPATIAL_STRIDES [[1],[2],[2],[2]] (slowfast/config/defaults.py:164),
                STRIDE_1X1 False -> stride sits on conv `b` (resnet_helper.py:265)
  * pool        pathway0_pool MaxPool3d [2,1,1] between s2 and s3 (:474-480, _POOL1 :76)
  * head        AvgPool3d [T/2, S/32, S/32] + Linear(2048->1) (:548-556, head_helper.py:74-95)
"""
from dataclasses import dataclass
from typing import List, Optional, Tuple

BN_EPS = 1e-5  # stem_helper.py:23, resnet_helper.py:212

STAGE_DEPTH = (3, 4, 6, 3)
STAGE_WIDTH_INNER = (64, 128, 256, 512)
STAGE_WIDTH_OUT = (256, 512, 1024, 2048)
STAGE_STRIDE = (1, 2, 2, 2)
# per-block temporal kernel of conv `a` after the reference's expansion
STAGE_TEMP_KERNELS = ((3, 3, 3), (3, 1, 3, 1), (3, 1, 3, 1, 3, 1), (1, 3, 1))
STEM_WIDTH = 64
FEATURE_DIM = 2048


@dataclass(frozen=True)
class ConvSpec:
    """One Conv3d(+eval BatchNorm3d) of the reference network."""
    name: str            # state_dict prefix of the conv weight, e.g. "resnet.s2.pathway0_res0.branch2.a"
    bn: str              # state_dict prefix of its BatchNorm3d
    cin: int
    cout: int
    kernel: Tuple[int, int, int]   # (kt, kh, kw)
    stride: Tuple[int, int, int]
    pad: Tuple[int, int, int]
    relu: bool           # ReLU applied right after BN (before any residual)
    pool2: bool = False  # FTCN-TT: MaxPool3d((1,2,2)) between this conv's BN and what follows


@dataclass(frozen=True)
class BlockSpec:
    stage: int           # 2..5
    index: int
    a: ConvSpec
    b: ConvSpec
    c: ConvSpec
    branch1: Optional[ConvSpec]


def stem_spec() -> ConvSpec:
    p = "resnet.s1.pathway0_stem"
    return ConvSpec(p + ".conv", p + ".bn", 3, STEM_WIDTH, (5, 7, 7), (1, 2, 2), (2, 3, 3), True)


def block_specs() -> List[BlockSpec]:
    out = []
    dim_in = STEM_WIDTH
    for si in range(4):
        stage = si + 2
        inner, dim_out, stride = STAGE_WIDTH_INNER[si], STAGE_WIDTH_OUT[si], STAGE_STRIDE[si]
        for bi in range(STAGE_DEPTH[si]):
            p = "resnet.s%d.pathway0_res%d" % (stage, bi)
            kt = STAGE_TEMP_KERNELS[si][bi]
            s = stride if bi == 0 else 1
            cin = dim_in if bi == 0 else dim_out
            a = ConvSpec(p + ".branch2.a", p + ".branch2.a_bn", cin, inner, (kt, 1, 1), (1, 1, 1), (kt // 2, 0, 0), True)
            b = ConvSpec(p + ".branch2.b", p + ".branch2.b_bn", inner, inner, (1, 3, 3), (1, s, s), (0, 1, 1), True)
            c = ConvSpec(p + ".branch2.c", p + ".branch2.c_bn", inner, dim_out, (1, 1, 1), (1, 1, 1), (0, 0, 0), False)
            br = None
            if cin != dim_out or s != 1:
                br = ConvSpec(p + ".branch1", p + ".branch1_bn", cin, dim_out, (1, 1, 1), (1, s, s), (0, 0, 0), False)
            out.append(BlockSpec(stage, bi, a, b, c, br))
        dim_in = dim_out
    return out


# ---------------------------------------------------------------------------------------------
# FTCN-TT variant (altfreezing/model/classifier/i3d_temporal_var_fix_dropout_tt_cfg.py with
# setting/ftcn_tt.yaml + root_setting.yaml: spatial_count 0, keep_stride_count 0, no_time_pool false,
# transformer.stop_point 5, patch_type time, depth 1, dim -1 -> 1024):
#   * temporal_only_conv (:207-289) rebuilds every conv whose spatial kernel is > 1 or whose spatial stride is 2
#     as a k[kt,1,1] s[1,1,1] conv; a removed stride becomes MaxPool3d((1,2,2)) appended to that conv's BatchNorm
#     (nn.Sequential(bn, pool) -> state_dict keys "...bn.0.*")
#   * s5 is replaced by nn.Identity (:315-321); the head is TransformerHead(spatial 14, time 16, 1024) (:323-330)
FTCN_STAGES = 3                      # s2..s4
TT_TOKENS, TT_DIM, TT_HEADS, TT_DIM_HEAD, TT_MLP, TT_DEPTH = 16, 1024, 16, 64, 2048, 1
TT_PREFIX = "resnet.head.time_T"

VARIANTS = ("i3d", "ftcn_tt")


def ftcn_stem_spec() -> ConvSpec:
    p = "resnet.s1.pathway0_stem"
    return ConvSpec(p + ".conv", p + ".bn.0", 3, STEM_WIDTH, (5, 1, 1), (1, 1, 1), (2, 0, 0), True, True)


def ftcn_block_specs() -> List[BlockSpec]:
    out = []
    for blk in block_specs():
        if blk.stage > FTCN_STAGES + 1:
            break
        strided = blk.b.stride[1] == 2
        b = ConvSpec(blk.b.name, blk.b.bn + (".0" if strided else ""), blk.b.cin, blk.b.cout, (1, 1, 1), (1, 1, 1),
                     (0, 0, 0), True, strided)
        br = blk.branch1
        if br is not None:
            br = ConvSpec(br.name, br.bn + (".0" if strided else ""), br.cin, br.cout, (1, 1, 1), (1, 1, 1), (0, 0, 0),
                          False, strided)
        out.append(BlockSpec(blk.stage, blk.index, blk.a, b, blk.c, br))
    return out


def tt_param_shapes():
    """state_dict name -> shape of the transformer head (TimeTransformer, time_transformer.py:219-279)."""
    inner = TT_HEADS * TT_DIM_HEAD
    shapes = {TT_PREFIX + ".pos_embedding": (1, TT_TOKENS + 1, TT_DIM), TT_PREFIX + ".cls_token": (1, 1, TT_DIM)}
    for i in range(TT_DEPTH):
        q = "%s.transformer.layers.%d" % (TT_PREFIX, i)
        shapes.update({
            q + ".0.fn.norm.weight": (TT_DIM,), q + ".0.fn.norm.bias": (TT_DIM,),
            q + ".0.fn.fn.to_qkv.weight": (3 * inner, TT_DIM),
            q + ".0.fn.fn.to_out.0.weight": (TT_DIM, inner), q + ".0.fn.fn.to_out.0.bias": (TT_DIM,),
            q + ".1.fn.norm.weight": (TT_DIM,), q + ".1.fn.norm.bias": (TT_DIM,),
            q + ".1.fn.fn.net.0.weight": (TT_MLP, TT_DIM), q + ".1.fn.fn.net.0.bias": (TT_MLP,),
            q + ".1.fn.fn.net.3.weight": (TT_DIM, TT_MLP), q + ".1.fn.fn.net.3.bias": (TT_DIM,),
        })
    shapes.update({TT_PREFIX + ".mlp_head.0.weight": (TT_DIM,), TT_PREFIX + ".mlp_head.0.bias": (TT_DIM,),
                   TT_PREFIX + ".mlp_head.1.weight": (1, TT_DIM), TT_PREFIX + ".mlp_head.1.bias": (1,)})
    return shapes


def stem_spec_for(variant: str) -> ConvSpec:
    return ftcn_stem_spec() if variant == "ftcn_tt" else stem_spec()


def block_specs_for(variant: str) -> List[BlockSpec]:
    return ftcn_block_specs() if variant == "ftcn_tt" else block_specs()


def feature_dim_for(variant: str) -> int:
    return TT_DIM if variant == "ftcn_tt" else FEATURE_DIM


def all_conv_specs(variant: str = "i3d") -> List[ConvSpec]:
    specs = [stem_spec_for(variant)]
    for blk in block_specs_for(variant):
        if blk.branch1 is not None:
            specs.append(blk.branch1)
        specs += [blk.a, blk.b, blk.c]
    return specs


def conv_out_dims(spec: ConvSpec, t: int, h: int, w: int) -> Tuple[int, int, int]:
    o = []
    for n, k, s, p in zip((t, h, w), spec.kernel, spec.stride, spec.pad):
        o.append((n + 2 * p - k) // s + 1)
    return tuple(o)


def macs_per_clip(t: int = 32, s: int = 224) -> int:
    """Multiply-accumulates of the 53 convolutions for one t x s x s clip
    (113.63 G for 32x224x224)."""
    total = 0
    st = stem_spec()
    d = conv_out_dims(st, t, s, s)
    total += d[0] * d[1] * d[2] * st.cout * st.cin * 5 * 7 * 7
    d = (d[0], (d[1] + 2 - 3) // 2 + 1, (d[2] + 2 - 3) // 2 + 1)
    for blk in block_specs():
        if blk.stage == 3 and blk.index == 0:
            d = (d[0] // 2, d[1], d[2])
        din = d
        for cv in (blk.a, blk.b, blk.c):
            d = conv_out_dims(cv, *d)
            total += d[0] * d[1] * d[2] * cv.cout * cv.cin * cv.kernel[0] * cv.kernel[1] * cv.kernel[2]
        if blk.branch1 is not None:
            e = conv_out_dims(blk.branch1, *din)
            total += e[0] * e[1] * e[2] * blk.branch1.cout * blk.branch1.cin
    return total
