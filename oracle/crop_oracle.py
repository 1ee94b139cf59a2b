"""TEST INFRASTRUCTURE ONLY — CPU oracle for the crop/align/normalise step.

numpy restatement of
  FasterCropAlignXRay.__call__/process_single  altfreezing/test_tools/faster_crop_align_xray.py:21-88
  estimiate_batch_transform & friends          altfreezing/test_tools/warp_for_xray.py:224-425,496-560
  get_crop_box                                 altfreezing/test_tools/utils.py:13-24
and of the arithmetic of the un-vendored dependency the reference calls,
`cv2.warpAffine(src u8C3, M, (S,S))` with default flags (INTER_LINEAR,
BORDER_CONSTANT 0).  OpenCV is not pinned by the reference
(altfreezing/requirements.txt lists no opencv); the build container has
opencv-python 4.13.0, whose u8 path is the legacy fixed-point remap
(coordinates quantised to 1/32 px, 15-bit weights).
Pinned against cv2 and the reference by tests/golden/make_golden.py.
Only tests/, smoke() and bench.py's CPU legs may import this file.
"""
import numpy as np

AB_BITS = 10
AB_SCALE = 1 << AB_BITS
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
INTER_REMAP_COEF_BITS = 15
INTER_REMAP_COEF_SCALE = 1 << INTER_REMAP_COEF_BITS

STD_POINTS_256 = np.array([           # warp_for_xray.py:532-549
    [85.82991, 115.7792], [169.0532, 114.3381], [127.574, 167.0006],
    [90.6964, 204.7014], [167.3069, 203.3733]]) + 30.0
STD_POINTS_256[:, 0] -= 30.0
STD_POINTS_256[:, 1] -= 60.0


def get_crop_box(shape, box, scale=0.5):
    """altfreezing/test_tools/utils.py:13-24"""
    height, width = shape
    box = np.rint(box).astype(int)
    nb = box.reshape(2, 2)
    size = nb[1] - nb[0]
    diff = (scale * size)[None, :] * np.array([-1, 1])[:, None]
    nb = nb + diff
    nb[:, 0] = np.clip(nb[:, 0], 0, width - 1)
    nb[:, 1] = np.clip(nb[:, 1], 0, height - 1)
    return np.rint(nb).astype(int).reshape(-1)


def _nonreflective(uv, xy):
    """warp_for_xray.py:224-334 (returns T only; Tinv is unused by the crop path)."""
    m = xy.shape[0]
    x = xy[:, 0].reshape(-1, 1)
    y = xy[:, 1].reshape(-1, 1)
    X = np.vstack((np.hstack((x, y, np.ones((m, 1)), np.zeros((m, 1)))),
                   np.hstack((y, -x, np.zeros((m, 1)), np.ones((m, 1))))))
    U = np.vstack((uv[:, 0].reshape(-1, 1), uv[:, 1].reshape(-1, 1)))
    if np.linalg.matrix_rank(X) < 4:
        raise Exception("cp2tform:twoUniquePointsReq")
    r = np.squeeze(np.linalg.lstsq(X, U, rcond=-1)[0])
    sc, ss, tx, ty = r[0], r[1], r[2], r[3]
    T = np.linalg.inv(np.array([[sc, -ss, 0], [ss, sc, 0], [tx, ty, 1]]))
    T[:, 2] = np.array([0, 0, 1])
    return T


def _tformfwd(trans, uv):
    return np.dot(np.hstack((uv, np.ones((uv.shape[0], 1)))), trans)[:, :2]


def estimate_batch_transform(all_src_pts, tgt_pts):
    """warp_for_xray.py:556-560 -> :496-529 -> findSimilarity :337-425.
    Reproduces the in-place aliasing at :404-405: `xyR = xy` is NOT a copy, so
    both residual norms are taken against the REFLECTED targets."""
    xy = np.repeat(np.asarray(tgt_pts, np.float64)[None], len(all_src_pts), 0).reshape(-1, 2)
    uv = np.array(all_src_pts, np.float64).reshape(-1, 2)
    trans1 = _nonreflective(uv, xy)
    xy[:, 0] = -1 * xy[:, 0]                      # mutates the array norm1/norm2 use
    trans2 = np.dot(_nonreflective(uv, xy), np.array([[-1, 0, 0], [0, 1, 0], [0, 0, 1]]))
    n1 = np.linalg.norm(_tformfwd(trans1, uv) - xy)
    n2 = np.linalg.norm(_tformfwd(trans2, uv) - xy)
    trans = trans1 if n1 <= n2 else trans2
    return trans[:, 0:2].T.copy(), trans


def _cv_round(v):
    return np.rint(v).astype(np.int64)            # cvRound: round-half-even


def bilinear_weight_table():
    """OpenCV's fixed-point bilinear table (imgwarp.cpp initInterTab2D, BilinearTab_i):
    int [32*32, 4], w = saturate_cast<short>(float32(wy*wx) * 32768).  Every entry sums to
    32768 except (fy,fx)=(0,0), where 32768 saturates to 32767; OpenCV then moves the
    missing 1 into another tap of that entry, which cannot change any u8 result
    ((32767*p + 16384) >> 15 == p for p <= 255) — checked against cv2 for all three
    plausible fix-up rules in tests/golden/make_golden.py."""
    tab1 = np.empty((INTER_TAB_SIZE, 2), np.float32)
    for i in range(INTER_TAB_SIZE):
        x = np.float32(i) * np.float32(1.0 / INTER_TAB_SIZE)
        tab1[i, 0] = np.float32(1.0) - x
        tab1[i, 1] = x
    out = np.empty((INTER_TAB_SIZE * INTER_TAB_SIZE, 4), np.int32)
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            v = np.array([tab1[i, k1] * tab1[j, k2] for k1 in range(2) for k2 in range(2)], np.float32)
            out[i * INTER_TAB_SIZE + j] = np.clip(np.rint(v.astype(np.float64) * INTER_REMAP_COEF_SCALE), -32768, 32767)
    return out


_WTAB = None


def invert_affine(M):
    """cv::invertAffineTransform in f64 (imgwarp.cpp), the first step of warpAffine."""
    M = np.asarray(M, np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    A12, A21 = -M[0, 1] * D, -M[1, 0] * D
    b1 = -A11 * M[0, 2] - A12 * M[1, 2]
    b2 = -A21 * M[0, 2] - A22 * M[1, 2]
    return np.array([[A11, A12, b1], [A21, A22, b2]], np.float64)


def warp_affine_u8(src, M, size, origin=(0, 0), valid_box=None, canvas_wh=None):
    """Integer-exact emulation of cv2.warpAffine(canvas, M, (size,size)) for u8 HxWxC.

    With the defaults `src` IS the canvas.  With `origin=(ox,oy)`, `valid_box`
    (x1,y1,x2,y2 in src coordinates, exclusive upper) and `canvas_wh`, `src` is the full decoded
    frame and canvas pixel (x,y) is frame pixel (x+ox, y+oy) masked to the frame's own
    big box — the zero-copy formulation the fused kernel uses."""
    global _WTAB
    if _WTAB is None:
        _WTAB = bilinear_weight_table().astype(np.int64)
    H, W = src.shape[:2]
    ox, oy = int(origin[0]), int(origin[1])
    iM = invert_affine(M)
    xs = np.arange(size, dtype=np.float64)
    adelta = _cv_round(iM[0, 0] * xs * AB_SCALE)
    bdelta = _cv_round(iM[1, 0] * xs * AB_SCALE)
    rd = AB_SCALE // INTER_TAB_SIZE // 2
    X0 = _cv_round((iM[0, 1] * xs + iM[0, 2]) * AB_SCALE) + rd
    Y0 = _cv_round((iM[1, 1] * xs + iM[1, 2]) * AB_SCALE) + rd
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    a = (Y & (INTER_TAB_SIZE - 1)) * INTER_TAB_SIZE + (X & (INTER_TAB_SIZE - 1))
    w = _WTAB[a]                                  # [S,S,4]
    if valid_box is None:
        x1, y1, x2, y2 = 0, 0, W, H
    else:
        x1, y1, x2, y2 = [int(v) for v in valid_box]
        x1, y1, x2, y2 = max(x1, 0), max(y1, 0), min(x2, W), min(y2, H)
    cw, ch = (W - ox, H - oy) if canvas_wh is None else canvas_wh

    def tap(yy, xx):
        fx, fy = xx + ox, yy + oy
        ok = (xx >= 0) & (xx < cw) & (yy >= 0) & (yy < ch) & (fx >= x1) & (fx < x2) & (fy >= y1) & (fy < y2)
        v = src[np.clip(fy, 0, H - 1), np.clip(fx, 0, W - 1)].astype(np.int64)
        return v * ok[..., None]

    acc = (tap(sy, sx) * w[..., 0:1] + tap(sy, sx + 1) * w[..., 1:2]
           + tap(sy + 1, sx) * w[..., 2:3] + tap(sy + 1, sx + 1) * w[..., 3:4])
    out = (acc + (1 << (INTER_REMAP_COEF_BITS - 1))) >> INTER_REMAP_COEF_BITS
    return np.clip(out, 0, 255).astype(np.uint8)


def clip_geometry(big_boxes, lm5_rel, size=224):
    """The per-clip quantities of FasterCropAlignXRay.__call__ (:22-50):
    left_top, canvas (w,h), per-frame diff, tfm (2x3) and trans (3x3)."""
    ori_boxes = np.asarray(big_boxes)
    left_top = ori_boxes[:, :2].min(0)
    right_bottom = ori_boxes[:, 2:].max(0)
    w, h = right_bottom - left_top
    diff = ori_boxes[:, :2] - left_top[None]
    new5 = np.asarray(lm5_rel) + diff[:, None, :]
    tfm, trans = estimate_batch_transform(new5.copy(), STD_POINTS_256 * size / 256.0)
    return left_top, (int(w), int(h)), diff, tfm, trans


def crop_align(landmarks, images, size=224):
    """FasterCropAlignXRay(size)(landmarks, images) -> (lm68_T [T,68,2], u8 [T,S,S,3]).
    landmarks: list of (box, lm5, lm68, big_box); images: list of HxWx3 u8 crops."""
    lm68 = np.array([lm[2] for lm in landmarks])
    left_top, (w, h), diff, tfm, trans = clip_geometry([lm[3] for lm in landmarks],
                                                       [lm[1] for lm in landmarks], size)
    new68 = lm68 + diff[:, None, :]
    lm68_t = np.array([np.dot(np.hstack((l, np.ones((l.shape[0], 1)))), trans)[:, :2] for l in new68])
    outs = []
    for img, d in zip(images, diff):
        canvas = np.zeros((h, w, 3), np.uint8)
        x, y = d
        ih, iw = img.shape[:2]
        canvas[y:y + ih, x:x + iw] = img
        outs.append(warp_affine_u8(canvas, tfm, size))
    return lm68_t, np.stack(outs)


def crop_align_from_frames(frames, big_boxes, tfm, left_top, canvas_wh, size=224):
    """Same pixels gathered straight from the decoded frames (no canvas copy)."""
    return np.stack([warp_affine_u8(f, tfm, size, origin=left_top, valid_box=bb, canvas_wh=canvas_wh)
                     for f, bb in zip(frames, big_boxes)])


def normalise(u8_thwc):
    """A4: x = (float(u8) - 255*mean_c) / (255*std_c), THWC -> CTHW, fp32
    (altfreezing/demo.py:84-87,317-319)."""
    mean = np.array([0.485, 0.456, 0.406], np.float32) * np.float32(255.0)
    std = np.array([0.229, 0.224, 0.225], np.float32) * np.float32(255.0)
    x = (u8_thwc.astype(np.float32) - mean) / std
    return np.ascontiguousarray(np.moveaxis(x, -1, 0))
