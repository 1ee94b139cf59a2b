"""ctypes binding of libafb200.so (C ABI: include/afb200.h).

There is deliberately no fallback: if the library has not been built (or does not
load) every entry raises.  Build it with `python __graft_entry__.py` or
`make -C <package>/csrc`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libafb200.so")

AF_OK = 0
AF_F32, AF_BF16, AF_F16, AF_U8 = 0, 1, 2, 3
AF_PREC_FP32, AF_PREC_BF16, AF_PREC_TF32 = 0, 1, 2


class AfConvDesc(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p),
                ("cin", C.c_int32), ("cout", C.c_int32),
                ("kt", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
                ("st", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
                ("pt", C.c_int32), ("ph", C.c_int32), ("pw", C.c_int32)]


class AfBlockDesc(C.Structure):
    _fields_ = [("branch1", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("c", C.c_int32),
                ("temporal_pool_before", C.c_int32), ("spatial_pool", C.c_int32)]


class AfTTLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ln1_w", "ln1_b", "qkv_w", "out_w", "out_b", "ln2_w", "ln2_b",
                                          "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class AfTTHead(C.Structure):
    _fields_ = [("dim", C.c_int32), ("tokens", C.c_int32), ("heads", C.c_int32), ("dim_head", C.c_int32),
                ("mlp_dim", C.c_int32), ("depth", C.c_int32), ("layers", C.POINTER(AfTTLayer)),
                ("cls_token", C.c_void_p), ("pos_embedding", C.c_void_p), ("norm_w", C.c_void_p),
                ("norm_b", C.c_void_p), ("fc_w", C.c_void_p), ("fc_b", C.c_float)]


class AfWeights(C.Structure):
    _fields_ = [("n_convs", C.c_int32), ("convs", C.POINTER(AfConvDesc)),
                ("stem", C.c_int32), ("n_blocks", C.c_int32), ("blocks", C.POINTER(AfBlockDesc)),
                ("fc_weight", C.c_void_p), ("fc_bias", C.c_float), ("feature_dim", C.c_int32),
                ("clip_t", C.c_int32), ("clip_s", C.c_int32),
                ("stem_pool2", C.c_int32), ("tt_head", C.POINTER(AfTTHead))]


class AfFrameDesc(C.Structure):
    _fields_ = [("data", C.c_void_p), ("pitch", C.c_int64), ("height", C.c_int32), ("width", C.c_int32),
                ("box", C.c_int32 * 4)]


class AfClipGeom(C.Structure):
    _fields_ = [("tfm", C.c_double * 6), ("left_top", C.c_int32 * 2), ("canvas_wh", C.c_int32 * 2)]


EXPORTS = ("af_last_error", "af_version", "af_launch_count", "af_create", "af_destroy", "af_set_option",
           "af_forward", "af_forward_frames", "af_infer_u8", "af_infer_u8_host", "af_submit_u8_host", "af_wait", "af_crop_u8", "af_crop_infer",
           "af_conv_ndhwc", "af_conv_shortcut_ndhwc", "af_get_stage", "af_get_stat", "af_crop_pack", "af_stem_pool_ndhwc4", "af_set_global_option", "af_ring_put_rows", "af_conv_bc_fused_ndhwc",
           "af_conv_bc_fused_tpool_ndhwc", "af_ring_put_boxes")

_lib = None


class Afb200Error(RuntimeError):
    pass


def lib():
    """Load libafb200.so once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Afb200Error("libafb200.so not built at %s — run `python __graft_entry__.py` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32p = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    L.af_last_error.restype = C.c_char_p
    L.af_last_error.argtypes = []
    L.af_version.restype = i32
    L.af_launch_count.restype = i64
    L.af_launch_count.argtypes = [vp]
    L.af_create.restype = i32
    L.af_create.argtypes = [C.POINTER(vp), i32, C.POINTER(AfWeights), i32, i32]
    L.af_destroy.restype = i32
    L.af_destroy.argtypes = [vp]
    L.af_set_option.restype = i32
    L.af_set_option.argtypes = [vp, C.c_char_p, i64]
    L.af_forward.restype = i32
    L.af_forward.argtypes = [vp, vp, i32, C.POINTER(i64), i32, f32p, f32p, vp]
    L.af_forward_frames.restype = i32
    L.af_forward_frames.argtypes = [vp, vp, i32, C.POINTER(i64), i32, f32p, f32p, vp]
    L.af_infer_u8.restype = i32
    L.af_infer_u8.argtypes = [vp, vp, i32, C.POINTER(C.c_float), C.POINTER(C.c_float), f32p, f32p, f32p, vp]
    L.af_infer_u8_host.restype = i32
    L.af_infer_u8_host.argtypes = [vp, vp, i32, C.POINTER(C.c_float), C.POINTER(C.c_float), f32p, f32p, vp]
    L.af_submit_u8_host.restype = i32
    L.af_submit_u8_host.argtypes = [vp, vp, i32, C.POINTER(C.c_float), C.POINTER(C.c_float), vp, C.POINTER(i32)]
    L.af_wait.restype = i32
    L.af_wait.argtypes = [vp, i32, f32p, f32p]
    L.af_crop_u8.restype = i32
    L.af_crop_u8.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.af_crop_infer.restype = i32
    L.af_crop_infer.argtypes = [vp, vp, vp, i32, i32, C.POINTER(C.c_float), C.POINTER(C.c_float), f32p, f32p, f32p, vp]
    L.af_conv_ndhwc.restype = i32
    L.af_conv_ndhwc.argtypes = [vp, C.POINTER(AfConvDesc), vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    L.af_conv_shortcut_ndhwc.restype = i32
    L.af_conv_shortcut_ndhwc.argtypes = [vp, C.POINTER(AfConvDesc), vp, C.POINTER(AfConvDesc), vp, i32, i32, i32, i32, i32,
                                         i32, i32, vp]
    L.af_crop_pack.restype = i32
    L.af_crop_pack.argtypes = [vp, vp, i32, i32, i32, i32, C.POINTER(C.c_float), C.POINTER(C.c_float), vp, i32,
                               C.POINTER(i64), vp]
    L.af_stem_pool_ndhwc4.restype = i32
    L.af_stem_pool_ndhwc4.argtypes = [vp, C.POINTER(AfConvDesc), vp, i32, i32, i32, i32, vp]
    L.af_set_global_option.restype = i32
    L.af_set_global_option.argtypes = [C.c_char_p, i64]
    L.af_ring_put_rows.restype = i32
    L.af_ring_put_rows.argtypes = [vp, i64, i64, i32, vp, vp, vp, vp, vp]
    L.af_ring_put_boxes.restype = i32
    L.af_ring_put_boxes.argtypes = [vp, i64, i64, i32, vp, vp, vp, vp]
    L.af_conv_bc_fused_ndhwc.restype = i32
    L.af_conv_bc_fused_ndhwc.argtypes = [vp, C.POINTER(AfConvDesc), C.POINTER(AfConvDesc), vp, vp, C.POINTER(AfConvDesc), vp, i32, i32, i32,
                                         i32, vp]
    L.af_conv_bc_fused_tpool_ndhwc.restype = i32
    L.af_conv_bc_fused_tpool_ndhwc.argtypes = [vp, C.POINTER(AfConvDesc), C.POINTER(AfConvDesc), vp, vp, i32, i32, i32, i32, vp]
    L.af_get_stat.restype = i32
    L.af_get_stat.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double)]
    L.af_get_stage.restype = i32
    L.af_get_stage.argtypes = [vp, i32, f32p, i64, C.POINTER(i32), vp]
    _lib = L
    return L


def check(rc, what=""):
    if rc != AF_OK:
        msg = lib().af_last_error().decode("utf-8", "replace")
        raise Afb200Error("%s failed (af_status %d): %s" % (what or "libafb200 call", rc, msg))
