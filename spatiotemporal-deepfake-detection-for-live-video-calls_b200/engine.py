"""Python host of the engine: owns an af_handle and exposes the reference-shaped calls.

  Engine.forward(x)            <- clf(x)["final_output"]        (altfreezing/demo.py:323-328)
  Engine.infer_scores_u8(arr)  <- ClassifierSvc.infer_scores    (altfreezing/TEST2.py:151-204)
  Engine.crop_infer(...)       <- crop_align_func + pack + classifier (altfreezing/demo.py:309-328)

PyTorch is used for device memory and streams only; all arithmetic runs in libafb200.so.
"""
import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib, synthetic
from ._lib import AF_BF16, AF_F16, AF_F32, AF_PREC_BF16, AF_PREC_FP32, AF_PREC_TF32, check, lib
from .weights import FoldedWeights

_DTYPES = {torch.float32: AF_F32, torch.bfloat16: AF_BF16, torch.float16: AF_F16}


def mean_std_255(style: str = "demo"):
    """The callers' normalisation constants times 255.  "demo": float32(0.485*255)
    (altfreezing/demo.py:84-87); "svc": float32(0.485)*255 in fp32 (TEST2.py:147-148)."""
    m, s = synthetic.IMAGENET_MEAN, synthetic.IMAGENET_STD
    if style == "demo":
        return (np.array([v * 255 for v in m], np.float32), np.array([v * 255 for v in s], np.float32))
    return (np.array(m, np.float32) * np.float32(255), np.array(s, np.float32) * np.float32(255))


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Engine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device: int = 0, max_batch: int = 32,
                 precision: str = "bf16", clip_t: int = 32, clip_s: int = 224, variant: str = "i3d"):
        if not torch.cuda.is_available():
            raise _lib.Afb200Error("afb200.Engine needs a CUDA device (sm_100a); there is no CPU fallback")
        self._L = lib()
        self.device = torch.device("cuda", device)
        self.max_batch, self.clip_t, self.clip_s = max_batch, clip_t, clip_s
        self.precision = precision
        self._h = C.c_void_p()
        self.variant = variant
        fw = FoldedWeights(state_dict, clip_t, clip_s, variant)
        self.feature_dim = int(fw.struct.feature_dim)
        # "tf32": fp32 storage / accumulation, trunk convs on the tensor cores with TF32 operands (conv_tf32.cu)
        prec = {"bf16": AF_PREC_BF16, "fp32": AF_PREC_FP32, "tf32": AF_PREC_TF32}[precision]
        torch.cuda.init()
        with torch.cuda.device(self.device):
            check(self._L.af_create(C.byref(self._h), device, C.byref(fw.struct), max_batch, prec), "af_create")
        self.mean255, self.std255 = mean_std_255("demo")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.af_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: int):
        check(self._L.af_set_option(self._h, name.encode(), int(value)), "af_set_option(%s)" % name)

    def get_stat(self, name: str) -> float:
        v = C.c_double()
        check(self._L.af_get_stat(self._h, name.encode(), C.byref(v)), "af_get_stat(%s)" % name)
        return float(v.value)

    @property
    def launch_count(self) -> int:
        return int(self._L.af_launch_count(self._h))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def forward(self, x: torch.Tensor, return_features: bool = False):
        """x: normalised clip tensor [B,3,T,S,S] on this engine's device (any strides;
        fp32 / bf16 / fp16).  Returns fp32 logits [B,1] (and features [B,feature_dim])."""
        if x.dim() != 5 or x.shape[1] != 3 or x.shape[2] != self.clip_t or x.shape[3] != self.clip_s \
                or x.shape[4] != self.clip_s:
            raise ValueError("afb200 engine takes [B,3,%d,%d,%d] clips, got %s" %
                             (self.clip_t, self.clip_s, self.clip_s, tuple(x.shape)))
        if x.device != self.device:
            raise ValueError("clip tensor is on %s, engine on %s" % (x.device, self.device))
        if x.dtype not in _DTYPES:
            x = x.float()
        B = x.shape[0]
        logits = torch.empty((B, 1), dtype=torch.float32, device=self.device)
        feats = torch.empty((B, self.feature_dim), dtype=torch.float32, device=self.device) if return_features else None
        strides = (C.c_int64 * 5)(*x.stride())
        with torch.cuda.device(self.device):
            for b0 in range(0, B, self.max_batch):
                nb = min(self.max_batch, B - b0)
                xb = x[b0:b0 + nb]
                check(self._L.af_forward(self._h, C.c_void_p(xb.data_ptr()), _DTYPES[x.dtype], strides, nb,
                                         C.c_void_p(logits[b0:].data_ptr()),
                                         C.c_void_p(feats[b0:].data_ptr()) if feats is not None else None,
                                         self._stream()), "af_forward")
        return (logits, feats) if return_features else logits

    def forward_frames(self, x: torch.Tensor):
        """x as in forward() -> (logits [B,1], frame features [B, T/2, 2048]): per-frame spatial means of the last
        stage, the `backbone(x) -> [B,T',D]` contract of dualrun's AltFreezingRGBEncoder."""
        if x.dim() != 5 or tuple(x.shape[1:]) != (3, self.clip_t, self.clip_s, self.clip_s):
            raise ValueError("afb200 engine takes [B,3,%d,%d,%d] clips, got %s" % (self.clip_t, self.clip_s, self.clip_s, tuple(x.shape)))
        if x.device != self.device:
            raise ValueError("clip tensor is on %s, engine on %s" % (x.device, self.device))
        if x.dtype not in _DTYPES:
            x = x.float()
        B = x.shape[0]
        logits = torch.empty((B, 1), dtype=torch.float32, device=self.device)
        feats = torch.empty((B, self.clip_t // 2, self.feature_dim), dtype=torch.float32, device=self.device)
        strides = (C.c_int64 * 5)(*x.stride())
        with torch.cuda.device(self.device):
            for b0 in range(0, B, self.max_batch):
                nb = min(self.max_batch, B - b0)
                check(self._L.af_forward_frames(self._h, C.c_void_p(x[b0:b0 + nb].data_ptr()), _DTYPES[x.dtype], strides, nb,
                                                C.c_void_p(logits[b0:].data_ptr()), C.c_void_p(feats[b0:].data_ptr()),
                                                self._stream()), "af_forward_frames")
        return logits, feats

    def infer_u8(self, clips: torch.Tensor, return_features: bool = False):
        """clips: u8 [B,T,S,S,3] RGB on the device -> (logits [B], scores [B])."""
        assert clips.dtype == torch.uint8 and clips.is_contiguous() and clips.device == self.device
        B = clips.shape[0]
        logits = torch.empty(B, dtype=torch.float32, device=self.device)
        scores = torch.empty(B, dtype=torch.float32, device=self.device)
        feats = torch.empty((B, self.feature_dim), dtype=torch.float32, device=self.device) if return_features else None
        with torch.cuda.device(self.device):
            for b0 in range(0, B, self.max_batch):
                nb = min(self.max_batch, B - b0)
                check(self._L.af_infer_u8(self._h, C.c_void_p(clips[b0:].data_ptr()), nb, _fptr(self.mean255),
                                          _fptr(self.std255), C.c_void_p(logits[b0:].data_ptr()),
                                          C.c_void_p(scores[b0:].data_ptr()),
                                          C.c_void_p(feats[b0:].data_ptr()) if feats is not None else None,
                                          self._stream()), "af_infer_u8")
        return (logits, scores, feats) if return_features else (logits, scores)

    def infer_scores_u8_host(self, aligned_batch_bthwc: np.ndarray, return_logits: bool = False):
        """ClassifierSvc.infer_scores: host u8 [B,T,S,S,3] -> host float32 scores [B]
        (H2D, pack, trunk, sigmoid, D2H inside the C call)."""
        arr = np.ascontiguousarray(aligned_batch_bthwc, dtype=np.uint8)
        B = arr.shape[0]
        logits = np.empty(B, np.float32)
        scores = np.empty(B, np.float32)
        with torch.cuda.device(self.device):
            for b0 in range(0, B, self.max_batch):
                nb = min(self.max_batch, B - b0)
                check(self._L.af_infer_u8_host(self._h, C.c_void_p(arr[b0:].ctypes.data), nb, _fptr(self.mean255),
                                               _fptr(self.std255), C.c_void_p(logits[b0:].ctypes.data),
                                               C.c_void_p(scores[b0:].ctypes.data), self._stream()),
                      "af_infer_u8_host")
        return (scores, logits) if return_logits else scores

    def infer_u8_host_ptr(self, host_ptr: int, batch: int, logits_ptr: int, scores_ptr: int):
        """Raw-pointer form of infer_scores_u8_host for pinned buffers (benchmark e2e leg)."""
        with torch.cuda.device(self.device):
            check(self._L.af_infer_u8_host(self._h, C.c_void_p(host_ptr), batch, _fptr(self.mean255),
                                           _fptr(self.std255), C.c_void_p(logits_ptr), C.c_void_p(scores_ptr),
                                           self._stream()), "af_infer_u8_host")

    def submit_u8_host_ptr(self, host_ptr: int, batch: int) -> int:
        """Pipelined scoring (af_submit_u8_host): queue upload + pack + trunk + read-back of `batch` u8 clips at
        `host_ptr` (pinned memory) and return a ticket at once; at most two tickets may be outstanding."""
        t = C.c_int32(-1)
        with torch.cuda.device(self.device):
            check(self._L.af_submit_u8_host(self._h, C.c_void_p(host_ptr), batch, _fptr(self.mean255), _fptr(self.std255),
                                            self._stream(), C.byref(t)), "af_submit_u8_host")
        return int(t.value)

    def wait(self, ticket: int, batch: int):
        """Block until `ticket`'s results are on the host -> (scores, logits) float32 [batch]."""
        logits = np.empty(batch, np.float32)
        scores = np.empty(batch, np.float32)
        with torch.cuda.device(self.device):
            check(self._L.af_wait(self._h, int(ticket), C.c_void_p(logits.ctypes.data), C.c_void_p(scores.ctypes.data)),
                  "af_wait")
        return scores, logits

    def crop_infer(self, frames_dev: torch.Tensor, geom_dev: torch.Tensor, batch: int, bgr: bool = False,
                   return_features: bool = False):
        """frames_dev: u8 tensor holding af_frame_desc[batch*T]; geom_dev: af_clip_geom[batch]
        (see crop.pack_descriptors).  Returns (logits [B], scores [B])."""
        logits = torch.empty(batch, dtype=torch.float32, device=self.device)
        scores = torch.empty(batch, dtype=torch.float32, device=self.device)
        feats = torch.empty((batch, self.feature_dim), dtype=torch.float32, device=self.device) if return_features else None
        with torch.cuda.device(self.device):
            check(self._L.af_crop_infer(self._h, C.c_void_p(frames_dev.data_ptr()), C.c_void_p(geom_dev.data_ptr()),
                                        batch, int(bgr), _fptr(self.mean255), _fptr(self.std255),
                                        C.c_void_p(logits.data_ptr()), C.c_void_p(scores.data_ptr()),
                                        C.c_void_p(feats.data_ptr()) if feats is not None else None,
                                        self._stream()), "af_crop_infer")
        return (logits, scores, feats) if return_features else (logits, scores)

    def get_stage(self, which: int) -> torch.Tensor:
        """fp32 NCTHW copy of stage `which` (1..5) of the last forward (needs keep_stages=1)."""
        dims = (C.c_int32 * 5)()
        check(self._L.af_get_stage(self._h, which, None, 0, dims, self._stream()), "af_get_stage")
        out = torch.empty(tuple(dims), dtype=torch.float32, device=self.device)
        check(self._L.af_get_stage(self._h, which, C.c_void_p(out.data_ptr()), out.numel(), dims, self._stream()),
              "af_get_stage")
        return out


def conv_ndhwc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, stride, pad, relu: bool,
               residual: Optional[torch.Tensor] = None, impl: int = 0, tf32: bool = False) -> torch.Tensor:
    """One folded conv on an NDHWC device tensor [B,T,H,W,C] (fp32 or bf16) through the
    engine's kernels (test/diagnostic entry af_conv_ndhwc). weight: [cout,cin,kt,kh,kw] fp32.
    tf32 (fp32 tensors only): the tensor-core kernel with TF32 operands instead of the exact FFMA kernel."""
    L = lib()
    assert x.is_cuda and x.is_contiguous() and x.dtype in (torch.float32, torch.bfloat16)
    B, T, H, W, Cin = x.shape
    w = weight.detach().float().cpu().contiguous().numpy()
    b = bias.detach().float().cpu().contiguous().numpy()
    cout, cin, kt, kh, kw = w.shape
    assert cin == Cin
    d = _lib.AfConvDesc()
    d.weight, d.bias = w.ctypes.data, b.ctypes.data
    d.cin, d.cout, d.kt, d.kh, d.kw = cin, cout, kt, kh, kw
    d.st, d.sh, d.sw = stride
    d.pt, d.ph, d.pw = pad
    To, Ho, Wo = [(n + 2 * p - k) // s + 1 for n, k, s, p in zip((T, H, W), (kt, kh, kw), stride, pad)]
    if impl == 4:                       # row-halo kernel with the fused 3x3/2 max-pool epilogue
        Ho, Wo = Ho // 2, Wo // 2
    if impl == 5:                       # pointwise kernel with the fused temporal [2,1,1] max-pool epilogue
        To = To // 2
    y = torch.empty((B, To, Ho, Wo, cout), dtype=x.dtype, device=x.device)
    if residual is not None:
        assert residual.dtype == x.dtype and residual.is_contiguous()
    prec = AF_PREC_BF16 if x.dtype == torch.bfloat16 else (AF_PREC_TF32 if tf32 else AF_PREC_FP32)
    with torch.cuda.device(x.device):
        check(L.af_conv_ndhwc(C.c_void_p(x.data_ptr()), C.byref(d),
                              C.c_void_p(residual.data_ptr()) if residual is not None else None,
                              C.c_void_p(y.data_ptr()), B, T, H, W, int(relu), prec, impl,
                              C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "af_conv_ndhwc")
    return y


def conv_bc_fused_ndhwc(x: torch.Tensor, weight_b, bias_b, weight_c, bias_c, residual: Optional[torch.Tensor] = None,
                        x2: Optional[torch.Tensor] = None, weight_s=None, bias_s=None, pool_t: bool = False) -> torch.Tensor:
    """relu(c(relu(b(x))) + residual) — or + shortcut(x2) — in ONE kernel (test/diagnostic entry af_conv_bc_fused_ndhwc):
    x bf16 [B,T,H,W,64], weight_b [64,64,1,3,3], weight_c [256,64,1,1,1], residual bf16 [B,T,H,W,256] or
    x2 bf16 [B,T,H,W,64] with weight_s [256,64,1,1,1].  pool_t (residual form only, af_conv_bc_fused_tpool_ndhwc): the
    max over frame pairs is taken in the epilogue and y is [B,T/2,H,W,256]."""
    assert (residual is None) != (x2 is None)
    if pool_t:
        assert residual is not None and x.shape[1] % 2 == 0
        for t in (x, residual):
            assert t.is_cuda and t.is_contiguous() and t.dtype == torch.bfloat16
        B, T, H, W, _ = x.shape
        db, keep_b = _conv_desc(weight_b, bias_b, (1, 1, 1), (0, 1, 1))
        dc, keep_c = _conv_desc(weight_c, bias_c, (1, 1, 1), (0, 0, 0))
        y = torch.empty((B, T // 2, H, W, dc.cout), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            check(lib().af_conv_bc_fused_tpool_ndhwc(C.c_void_p(x.data_ptr()), C.byref(db), C.byref(dc),
                                                     C.c_void_p(residual.data_ptr()), C.c_void_p(y.data_ptr()), B, T, H, W,
                                                     C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                  "af_conv_bc_fused_tpool_ndhwc")
        del keep_b, keep_c
        return y
    for t in (x, residual, x2):
        assert t is None or (t.is_cuda and t.is_contiguous() and t.dtype == torch.bfloat16)
    B, T, H, W, _ = x.shape
    db, keep_b = _conv_desc(weight_b, bias_b, (1, 1, 1), (0, 1, 1))
    dc, keep_c = _conv_desc(weight_c, bias_c, (1, 1, 1), (0, 0, 0))
    ds, keep_s = _conv_desc(weight_s, bias_s, (1, 1, 1), (0, 0, 0)) if x2 is not None else (None, None)
    y = torch.empty((B, T, H, W, dc.cout), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().af_conv_bc_fused_ndhwc(C.c_void_p(x.data_ptr()), C.byref(db), C.byref(dc),
                                           C.c_void_p(residual.data_ptr()) if residual is not None else None,
                                           C.c_void_p(x2.data_ptr()) if x2 is not None else None,
                                           C.byref(ds) if ds is not None else None, C.c_void_p(y.data_ptr()), B, T, H, W,
                                           C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "af_conv_bc_fused_ndhwc")
    del keep_b, keep_c, keep_s
    return y


def set_global_option(name: str, value: int):
    """Process-wide diagnostic knob (af_set_global_option), e.g. ("block_n", 64|128|256|0)."""
    check(lib().af_set_global_option(name.encode(), int(value)), "af_set_global_option(%s)" % name)


def stem_pool_ndhwc4(clip: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, per_frame_kernel: bool = False) -> torch.Tensor:
    """The fused stem kernel on its own (test/diagnostic entry af_stem_pool_ndhwc4): clip bf16 [B,T,S,S,4] (channel 3
    ignored), folded stem weight [64,3,5,7,7] + bias [64] -> relu(conv) max-pooled 3x3/2, bf16 [B,T,S/4,S/4,64]."""
    assert clip.is_cuda and clip.is_contiguous() and clip.dtype == torch.bfloat16 and clip.shape[-1] == 4
    B, T, S, S2, _ = clip.shape
    assert S == S2
    d, keep = _conv_desc(weight, bias, (1, 2, 2), (2, 3, 3))
    y = torch.empty((B, T, S // 4, S // 4, 64), dtype=torch.bfloat16, device=clip.device)
    with torch.cuda.device(clip.device):
        check(lib().af_stem_pool_ndhwc4(C.c_void_p(clip.data_ptr()), C.byref(d), C.c_void_p(y.data_ptr()), B, T, S,
                                        int(per_frame_kernel), C.c_void_p(torch.cuda.current_stream(clip.device).cuda_stream)),
              "af_stem_pool_ndhwc4")
    del keep
    return y


def _conv_desc(weight: torch.Tensor, bias: torch.Tensor, stride, pad):
    w = weight.detach().float().cpu().contiguous().numpy()
    b = bias.detach().float().cpu().contiguous().numpy()
    d = _lib.AfConvDesc()
    d.weight, d.bias = w.ctypes.data, b.ctypes.data
    d.cout, d.cin, d.kt, d.kh, d.kw = w.shape
    d.st, d.sh, d.sw = stride
    d.pt, d.ph, d.pw = pad
    return d, (w, b)


def conv_shortcut_ndhwc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, x2: torch.Tensor,
                        weight2: torch.Tensor, bias2: torch.Tensor, stride2, relu: bool = True) -> torch.Tensor:
    """relu(conv1x1x1(x) + conv1x1x1_stride(x2) + both biases) in ONE tcgen05 GEMM — the first ResBlock of a
    stage with its projection shortcut fused into the `c` conv (test/diagnostic entry af_conv_shortcut_ndhwc).
    x: bf16 [B,T,Ho,Wo,cin]; x2: bf16 [B,T,H2,W2,cin2] with (H2-1)//sh+1 == Ho; weights [cout,cin,1,1,1]."""
    L = lib()
    for t in (x, x2):
        assert t.is_cuda and t.is_contiguous() and t.dtype == torch.bfloat16
    B, T, H, W, _ = x.shape
    d1, keep1 = _conv_desc(weight, bias, (1, 1, 1), (0, 0, 0))
    d2, keep2 = _conv_desc(weight2, bias2, stride2, (0, 0, 0))
    y = torch.empty((B, T, H, W, d1.cout), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(L.af_conv_shortcut_ndhwc(C.c_void_p(x.data_ptr()), C.byref(d1), C.c_void_p(x2.data_ptr()), C.byref(d2),
                                       C.c_void_p(y.data_ptr()), B, T, H, W, x2.shape[2], x2.shape[3], int(relu),
                                       C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
              "af_conv_shortcut_ndhwc")
    del keep1, keep2
    return y
