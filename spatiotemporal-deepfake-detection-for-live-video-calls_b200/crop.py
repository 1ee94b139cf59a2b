"""Host side of the crop/align step: the per-clip geometry (numpy, fp64) and the
FasterCropAlignXRay-compatible wrapper around the K1 CUDA kernel.

  get_crop_box                     <- altfreezing/test_tools/utils.py:13-24
  estimate_clip_transform          <- warp_for_xray.py:556-560 -> :496-529 -> findSimilarity :337-425
                                      -> findNonreflectiveSimilarity :224-334
  clip_geometry / CropAlignB200    <- FasterCropAlignXRay.__call__ faster_crop_align_xray.py:21-75

The pixels are produced on the GPU (csrc/crop_pack.cu); only the 4-dof least-squares fit
(a [2*5T,4] system, ~0.3 ms) and the landmark transforms stay on the host.
"""
import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np
import torch

from ._lib import AfClipGeom, AfFrameDesc, check, lib

_STD_317 = np.array([[85.82991, 115.7792], [169.0532, 114.3381], [127.574, 167.0006],
                     [90.6964, 204.7014], [167.3069, 203.3733]]) + 30.0
STD_POINTS_256 = _STD_317.copy()
STD_POINTS_256[:, 0] -= 30.0
STD_POINTS_256[:, 1] -= 60.0


def get_crop_box(shape, box, scale=0.5):
    """Enlarge a detector box by `scale` of its size on every side, clip to the frame,
    round to int -> (x1, y1, x2, y2)."""
    height, width = shape
    b = np.rint(np.asarray(box, np.float64)).astype(int).reshape(2, 2)
    half = scale * (b[1] - b[0])
    nb = b + np.stack([-half, half])
    nb[:, 0] = np.clip(nb[:, 0], 0, width - 1)
    nb[:, 1] = np.clip(nb[:, 1], 0, height - 1)
    return np.rint(nb).astype(int).reshape(-1)


def _fit_nonreflective(src, dst):
    """Least-squares non-reflective similarity: returns the 3x3 `T` with [dst 1] = [src 1] @ T
    (solves for the inverse map dst->src, then inverts it, as the reference does)."""
    n = dst.shape[0]
    x, y = dst[:, 0:1], dst[:, 1:2]
    one, zero = np.ones((n, 1)), np.zeros((n, 1))
    X = np.vstack((np.hstack((x, y, one, zero)), np.hstack((y, -x, zero, one))))
    U = np.vstack((src[:, 0:1], src[:, 1:2]))
    sol, _, rank, _ = np.linalg.lstsq(X, U, rcond=-1)      # same solver call as the reference; its rank replaces
    if rank < 4:                                           # the reference's separate matrix_rank() SVD
        raise ValueError("similarity fit needs at least two distinct points")
    r = np.squeeze(sol)
    T = np.linalg.inv(np.array([[r[0], -r[1], 0.0], [r[1], r[0], 0.0], [r[2], r[3], 1.0]]))
    T[:, 2] = (0.0, 0.0, 1.0)
    return T


def _apply(T, pts):
    return (np.hstack((pts, np.ones((pts.shape[0], 1)))) @ T)[:, :2]


def estimate_clip_transform(src_pts_t52, tgt_pts_52) -> Tuple[np.ndarray, np.ndarray]:
    """One similarity for the whole clip from all T x 5 landmark pairs -> (tfm 2x3, trans 3x3).

    Behavioural quirk kept on purpose: the reference reflects its target array IN PLACE
    before fitting the mirrored candidate (`xyR = xy` is an alias, warp_for_xray.py:404-405),
    so BOTH residuals are measured against the mirrored targets."""
    src = np.asarray(src_pts_t52, np.float64).reshape(-1, 2)
    dst = np.repeat(np.asarray(tgt_pts_52, np.float64)[None], len(src_pts_t52), 0).reshape(-1, 2)
    direct = _fit_nonreflective(src, dst)
    dst[:, 0] *= -1.0
    mirrored = _fit_nonreflective(src, dst) @ np.diag([-1.0, 1.0, 1.0])
    e_direct = np.linalg.norm(_apply(direct, src) - dst)
    e_mirrored = np.linalg.norm(_apply(mirrored, src) - dst)
    trans = direct if e_direct <= e_mirrored else mirrored
    return trans[:, 0:2].T.copy(), trans


def clip_geometry(big_boxes, lm5_rel, size=224):
    """left_top, canvas (w,h), per-frame offset, tfm, trans for one clip.
    big_boxes: [T,4] int frame coordinates; lm5_rel: [T,5,2] relative to each frame's big box."""
    boxes = np.asarray(big_boxes)
    left_top = boxes[:, :2].min(0)
    w, h = boxes[:, 2:].max(0) - left_top
    diff = boxes[:, :2] - left_top[None]
    lm5 = np.asarray(lm5_rel, np.float64) + diff[:, None, :]
    tfm, trans = estimate_clip_transform(lm5, STD_POINTS_256 * size / 256.0)
    return left_top, (int(w), int(h)), diff, tfm, trans


def get_crop_boxes(shape, boxes, scale=0.5):
    """get_crop_box for an [N,4] array of detector boxes in one vectorised pass (same arithmetic element by element,
    so the results are identical to N scalar calls) -> int [N,4]."""
    height, width = shape
    b = np.rint(np.asarray(boxes, np.float64)).astype(int).reshape(-1, 2, 2)
    half = scale * (b[:, 1] - b[:, 0])
    nb = b + np.stack([-half, half], axis=1)
    nb[:, :, 0] = np.clip(nb[:, :, 0], 0, width - 1)
    nb[:, :, 1] = np.clip(nb[:, :, 1], 0, height - 1)
    return np.rint(nb).astype(int).reshape(-1, 4)


def _fit_nonreflective_batch(src, dst):
    """_fit_nonreflective for B point sets sharing ONE target layout: src [B,n,2], dst [n,2] -> T [B,3,3].  The
    system matrix depends on dst only and is built once."""
    n = dst.shape[0]
    x, y = dst[:, 0:1], dst[:, 1:2]
    one, zero = np.ones((n, 1)), np.zeros((n, 1))
    X = np.vstack((np.hstack((x, y, one, zero)), np.hstack((y, -x, zero, one))))
    U = np.concatenate((src[:, :, 0], src[:, :, 1]), axis=1)              # [B, 2n]
    r = np.empty((src.shape[0], 4))
    for b in range(src.shape[0]):       # one single-right-hand-side solve per clip, exactly the scalar routine's call
        sol, _, rank, _ = np.linalg.lstsq(X, U[b][:, None], rcond=-1)    # (LAPACK's multi-RHS path is 1000x slower here)
        if rank < 4:
            raise ValueError("similarity fit needs at least two distinct points")
        r[b] = sol[:, 0]
    M = np.zeros((src.shape[0], 3, 3))
    M[:, 0, 0], M[:, 0, 1] = r[:, 0], -r[:, 1]
    M[:, 1, 0], M[:, 1, 1] = r[:, 1], r[:, 0]
    M[:, 2, 0], M[:, 2, 1], M[:, 2, 2] = r[:, 2], r[:, 3], 1.0
    T = np.linalg.inv(M)
    T[:, :, 2] = (0.0, 0.0, 1.0)
    return T


def clip_geometry_batch(big_boxes, lm5_rel, size=224):
    """clip_geometry for B clips at once (the host side of a 32-clip step costs one lstsq pair instead of 64):
    big_boxes [B,T,4] int, lm5_rel [B,T,5,2] -> list of (tfm 2x3, left_top (x,y), canvas (w,h)) per clip.
    Same algorithm and the same lstsq calls, incl. the reference's mirrored-target quirk; everything around the
    solves is vectorised over the clips (transforms agree with the per-clip routine to ~1e-12)."""
    boxes = np.asarray(big_boxes)
    B, T = boxes.shape[:2]
    left_top = boxes[:, :, :2].min(1)                                     # [B,2]
    wh = boxes[:, :, 2:].max(1) - left_top
    diff = boxes[:, :, :2] - left_top[:, None]
    src = (np.asarray(lm5_rel, np.float64) + diff[:, :, None, :]).reshape(B, T * 5, 2)
    dst = np.repeat((STD_POINTS_256 * size / 256.0)[None], T, 0).reshape(-1, 2)
    direct = _fit_nonreflective_batch(src, dst)
    dst_m = dst.copy()
    dst_m[:, 0] *= -1.0
    mirrored = _fit_nonreflective_batch(src, dst_m) @ np.diag([-1.0, 1.0, 1.0])
    src1 = np.concatenate((src, np.ones((B, T * 5, 1))), axis=2)
    e_d = np.linalg.norm(((src1 @ direct)[:, :, :2] - dst_m[None]).reshape(B, -1), axis=1)     # both vs the mirrored targets
    e_m = np.linalg.norm(((src1 @ mirrored)[:, :, :2] - dst_m[None]).reshape(B, -1), axis=1)
    trans = np.where((e_d <= e_m)[:, None, None], direct, mirrored)
    return [(trans[b][:, 0:2].T.copy(), left_top[b], (int(wh[b, 0]), int(wh[b, 1]))) for b in range(B)]


def pack_descriptors(frame_tensors: Sequence[torch.Tensor], big_boxes, geoms, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Build the device arrays af_crop_u8 / af_crop_infer take.
    frame_tensors: B*T u8 CUDA tensors [H,W,3] (may alias one another; may be row-strided views);
    big_boxes: [B*T,4]; geoms: list of (tfm 2x3, left_top (x,y), canvas (w,h)) per clip."""
    n = len(frame_tensors)
    fd = (AfFrameDesc * n)()
    for i, (ft, bb) in enumerate(zip(frame_tensors, big_boxes)):
        if isinstance(ft, torch.Tensor):
            assert ft.dtype == torch.uint8 and ft.dim() == 3 and ft.shape[2] == 3 and ft.stride(2) == 1 and ft.stride(1) == 3
            ptr, pitch, hh, ww = ft.data_ptr(), ft.stride(0), ft.shape[0], ft.shape[1]
        else:                       # raw (device pointer, pitch, height, width)
            ptr, pitch, hh, ww = ft
        fd[i].data = ptr
        fd[i].pitch = pitch
        fd[i].height, fd[i].width = hh, ww
        for k in range(4):
            fd[i].box[k] = int(bb[k])
    cg = (AfClipGeom * len(geoms))()
    for i, (tfm, lt, wh) in enumerate(geoms):
        flat = np.asarray(tfm, np.float64).reshape(-1)
        for k in range(6):
            cg[i].tfm[k] = float(flat[k])
        cg[i].left_top[0], cg[i].left_top[1] = int(lt[0]), int(lt[1])
        cg[i].canvas_wh[0], cg[i].canvas_wh[1] = int(wh[0]), int(wh[1])
    fd_t = torch.frombuffer(bytearray(bytes(fd)), dtype=torch.uint8).to(device)
    cg_t = torch.frombuffer(bytearray(bytes(cg)), dtype=torch.uint8).to(device)
    return fd_t, cg_t


FRAME_DESC_DTYPE = np.dtype([("data", "<u8"), ("pitch", "<i8"), ("height", "<i4"), ("width", "<i4"), ("box", "<i4", (4,))])
CLIP_GEOM_DTYPE = np.dtype([("tfm", "<f8", (6,)), ("left_top", "<i4", (2,)), ("canvas_wh", "<i4", (2,))])
assert FRAME_DESC_DTYPE.itemsize == 40 and CLIP_GEOM_DTYPE.itemsize == 64


def pack_descriptors_ring(base_ptr: int, slot_stride: int, pitch: int, height: int, width: int, slots, big_boxes, geoms,
                          device):
    """Vectorised descriptor build for frames living in ONE device ring (live path): frame i is
    `base_ptr + slots[i] * slot_stride`.  Same output as pack_descriptors, without a Python loop per frame."""
    n = len(slots)
    fd = np.zeros(n, FRAME_DESC_DTYPE)
    fd["data"] = np.uint64(base_ptr) + np.asarray(slots, np.uint64) * np.uint64(slot_stride)
    fd["pitch"], fd["height"], fd["width"] = pitch, height, width
    fd["box"] = np.asarray(big_boxes, np.int32).reshape(n, 4)
    cg = np.zeros(len(geoms), CLIP_GEOM_DTYPE)
    cg["tfm"] = np.stack([np.asarray(g[0], np.float64).reshape(6) for g in geoms])
    cg["left_top"] = np.stack([np.asarray(g[1], np.int32).reshape(2) for g in geoms])
    cg["canvas_wh"] = np.stack([np.asarray(g[2], np.int32).reshape(2) for g in geoms])
    fd_t = torch.from_numpy(fd.view(np.uint8)).to(device, non_blocking=True)
    cg_t = torch.from_numpy(cg.view(np.uint8)).to(device, non_blocking=True)
    return fd_t, cg_t


def crop_u8(frame_tensors, big_boxes, geoms, frames_per_clip, size=224, bgr=False, device=None) -> torch.Tensor:
    """GPU warp of B clips -> u8 [B,T,S,S,3] on the device (bit-exact with cv2.warpAffine)."""
    dev = device if device is not None else frame_tensors[0].device
    B = len(geoms)
    fd_t, cg_t = pack_descriptors(frame_tensors, big_boxes, geoms, dev)
    out = torch.empty((B, frames_per_clip, size, size, 3), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().af_crop_u8(C.c_void_p(fd_t.data_ptr()), C.c_void_p(cg_t.data_ptr()), B, frames_per_clip, size,
                               int(bgr), C.c_void_p(out.data_ptr()),
                               C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "af_crop_u8")
    return out


def crop_pack(frame_tensors, big_boxes, geoms, frames_per_clip, size, mean255, std255, out: torch.Tensor = None,
              dtype=torch.bfloat16, bgr=False, device=None) -> torch.Tensor:
    """K1 on its own (af_crop_pack): warp + `(u8 - 255*mean)/(255*std)` written straight into `out`, a [B,3,T,S,S]
    fp32/bf16 CUDA tensor with ANY strides (contiguous NCTHW, channels_last_3d, a permuted NTHWC buffer ...);
    allocated contiguous when not given.  Equals crop_u8 followed by the callers' pack lines bit for bit."""
    dev = device if device is not None else (out.device if out is not None else frame_tensors[0].device)
    B = len(geoms)
    if out is None:
        out = torch.empty((B, 3, frames_per_clip, size, size), dtype=dtype, device=dev)
    assert out.is_cuda and tuple(out.shape) == (B, 3, frames_per_clip, size, size) and out.dtype in (torch.float32, torch.bfloat16)
    fd_t, cg_t = pack_descriptors(frame_tensors, big_boxes, geoms, dev)
    m = np.ascontiguousarray(mean255, np.float32)
    sd = np.ascontiguousarray(std255, np.float32)
    strides = (C.c_int64 * 5)(*out.stride())
    with torch.cuda.device(dev):
        check(lib().af_crop_pack(C.c_void_p(fd_t.data_ptr()), C.c_void_p(cg_t.data_ptr()), B, frames_per_clip, size, int(bgr),
                                 m.ctypes.data_as(C.POINTER(C.c_float)), sd.ctypes.data_as(C.POINTER(C.c_float)),
                                 C.c_void_p(out.data_ptr()), 0 if out.dtype == torch.float32 else 1, strides,
                                 C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "af_crop_pack")
    return out


class CropAlignB200:
    """Drop-in for FasterCropAlignXRay(size): `(landmarks, images) -> (lm68_T, u8 [T,S,S,3])`.
    landmarks: list of (box, lm5, lm68, big_box) with lm coordinates relative to the big box;
    images: list of HxWx3 u8 crops (numpy; views are fine).  The crops are uploaded and warped
    by the CUDA kernel; results come back as numpy like the reference's."""

    def __init__(self, size=256, return_ldm5=False, device=0):
        self.image_size = size
        self.std_points = STD_POINTS_256 * size / 256.0
        self.return_ldm5 = return_ldm5
        self.device = torch.device("cuda", device)

    def __call__(self, landmarks, images=None, jitter=False):
        landmarks = [lm[:4] for lm in landmarks]
        boxes = np.array([lm[3] for lm in landmarks])
        lm5 = np.array([lm[1] for lm in landmarks])
        lm68 = np.array([lm[2] for lm in landmarks])
        left_top = boxes[:, :2].min(0)
        w, h = boxes[:, 2:].max(0) - left_top
        diff = boxes[:, :2] - left_top[None]
        new5 = lm5 + diff[:, None, :]
        new68 = lm68 + diff[:, None, :]
        fit = new5.copy()
        if jitter:
            fit += np.random.uniform(-4, 4, fit.shape)
        tfm, trans = estimate_clip_transform(fit, self.std_points)
        t68 = np.array([_apply(trans, l) for l in new68])
        t5 = np.array([_apply(trans, l) for l in new5])
        if images is None:
            return (t5, t68) if self.return_ldm5 else t68
        # Each crop is uploaded as is and addressed in CANVAS coordinates: the descriptor's
        # base pointer is moved back by the crop's offset d inside the canvas, so canvas pixel
        # (x,y) resolves to crop pixel (x-d.x, y-d.y); the box mask keeps every read inside the crop.
        # The kernel fetches interior taps as aligned 16-byte windows (crop_pack.cu: load6), which may reach up to
        # 15 bytes past either end of a crop: every crop sits in one upload buffer with 16 bytes of slack around it.
        keep, descs, bbs = [], [], []
        sizes = [int(img.shape[0]) * int(img.shape[1]) * 3 for img in images]
        offs = np.concatenate([[0], np.cumsum([(n + 16 + 15) // 16 * 16 for n in sizes])]).astype(np.int64) + 16
        host = np.zeros(int(offs[-1]) + 16, np.uint8)
        for img, o, n in zip(images, offs, sizes):
            host[o:o + n] = np.ascontiguousarray(img).reshape(-1)
        buf = torch.from_numpy(host).to(self.device)
        keep.append(buf)
        for img, d, o in zip(images, diff, offs):
            ih, iw = img.shape[:2]
            pitch = iw * 3
            descs.append((buf.data_ptr() + int(o) - (int(d[1]) * pitch + int(d[0]) * 3), pitch, int(h), int(w)))
            bbs.append((int(d[0]), int(d[1]), int(d[0]) + iw, int(d[1]) + ih))
        out = crop_u8(descs, bbs, [(tfm, (0, 0), (int(w), int(h)))], len(images), self.image_size,
                      device=self.device)
        imgs = out[0].cpu().numpy()
        return (t5, t68, imgs) if self.return_ldm5 else (t68, imgs)
