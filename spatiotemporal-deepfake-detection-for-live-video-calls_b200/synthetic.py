"""Seeded synthetic weights, clips, frames and face tracks.

No AltFreezing checkpoint ships with the reference (altfreezing/.gitignore:163)
and there is no network, so parity and the benchmark run on synthetic inputs of
the right shape.  The reference's own random init is degenerate for testing
(every block-final BN gamma is 0, slowfast/utils/weight_init_helper.py:29-36),
so BN statistics are re-randomised (SURVEY.md §8c recipe).

Everything here is deterministic in (seed) on the torch/numpy CPU generators.
"""
import math
from typing import Dict

import numpy as np
import torch

from . import arch

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # altfreezing/demo.py:84-87
IMAGENET_STD = (0.229, 0.224, 0.225)


def synthetic_state_dict(seed: int = 0, variant: str = "i3d") -> Dict[str, torch.Tensor]:
    """A full state_dict of the reference network (320 keys for the `I3D8x8` of i3d_ori; 275 for
    the FTCN-TT plugin) with MSRA(fan_out) conv weights, non-degenerate BN and, for FTCN-TT, a transformer head
    drawn like TimeTransformer._init_weights (normal 0.02 Linear weights) but with non-trivial biases/LayerNorms."""
    g = torch.Generator().manual_seed(int(seed))
    sd: Dict[str, torch.Tensor] = {}

    def randn(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32)

    def rand(lo, hi, *shape):
        return lo + (hi - lo) * torch.rand(*shape, generator=g, dtype=torch.float32)

    for cv in arch.all_conv_specs(variant):
        kt, kh, kw = cv.kernel
        fan_out = cv.cout * kt * kh * kw
        sd[cv.name + ".weight"] = randn(cv.cout, cv.cin, kt, kh, kw) * math.sqrt(2.0 / fan_out)
        final = cv.name.endswith("branch2.c")
        sd[cv.bn + ".weight"] = rand(0.2, 0.5, cv.cout) if final else rand(0.8, 1.2, cv.cout)
        sd[cv.bn + ".bias"] = randn(cv.cout) * 0.05
        sd[cv.bn + ".running_mean"] = randn(cv.cout) * 0.05
        sd[cv.bn + ".running_var"] = rand(0.8, 1.2, cv.cout)
        sd[cv.bn + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    if variant == "ftcn_tt":
        for name, shape in arch.tt_param_shapes().items():
            if name.endswith("norm.weight") or name.endswith("mlp_head.0.weight"):
                sd[name] = rand(0.8, 1.2, *shape)
            elif name.endswith(".bias"):
                sd[name] = randn(*shape) * 0.05
            elif name.endswith("pos_embedding") or name.endswith("cls_token"):
                sd[name] = randn(*shape) * 0.5
            else:                                   # Linear weights
                sd[name] = randn(*shape) * (0.05 if name.endswith("mlp_head.1.weight") else 0.03)
    else:
        sd["resnet.head.projection.weight"] = randn(1, arch.FEATURE_DIM) * 0.05
        sd["resnet.head.projection.bias"] = torch.full((1,), 0.1)
    return sd


def synthetic_clip_u8(index: int, t: int = 32, s: int = 224) -> np.ndarray:
    """One u8 [t,s,s,3] RGB face-crop clip: 5x5 box-filtered noise times a
    per-clip gain plus a moving Gaussian blob."""
    rng = np.random.default_rng(1000 + int(index))
    noise = rng.integers(0, 256, size=(t, s + 4, s + 4, 3)).astype(np.float32)
    c = np.cumsum(np.cumsum(noise, axis=1), axis=2)
    c = np.pad(c, ((0, 0), (1, 0), (1, 0), (0, 0)))
    box = (c[:, 5:, 5:] - c[:, :-5, 5:] - c[:, 5:, :-5] + c[:, :-5, :-5]) / 25.0
    gain = rng.uniform(0.5, 1.5)
    yy, xx = np.mgrid[0:s, 0:s].astype(np.float32)
    cx0, cy0 = rng.uniform(0.3 * s, 0.7 * s, size=2)
    vx, vy = rng.uniform(-1.5, 1.5, size=2)
    sig = rng.uniform(0.08 * s, 0.2 * s)
    amp = rng.uniform(40.0, 120.0) * rng.choice([-1.0, 1.0])
    out = np.empty((t, s, s, 3), np.float32)
    for f in range(t):
        blob = amp * np.exp(-((xx - cx0 - vx * f) ** 2 + (yy - cy0 - vy * f) ** 2) / (2 * sig * sig))
        out[f] = box[f] * gain + blob[..., None]
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def normalise_clip(u8_bthwc: np.ndarray) -> torch.Tensor:
    """The callers' pack step, `x = (u8 - 255*mean) / (255*std)` then
    NTHWC -> NCTHW (altfreezing/demo.py:317-319, TEST2.py:153-158)."""
    x = torch.as_tensor(u8_bthwc, dtype=torch.float32)
    if x.dim() == 4:
        x = x.unsqueeze(0)
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32) * 255.0
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32) * 255.0
    return x.permute(0, 4, 1, 2, 3).sub(mean.view(1, 3, 1, 1, 1)).div(std.view(1, 3, 1, 1, 1))


def synthetic_frame_u8(index: int, h: int = 720, w: int = 1280) -> np.ndarray:
    """One decoded u8 [h,w,3] frame (smooth gradients + noise), seed 2000+index."""
    rng = np.random.default_rng(2000 + int(index))
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    base = np.stack([
        127 + 100 * np.sin(xx / rng.uniform(20, 60) + rng.uniform(0, 6)) * np.cos(yy / rng.uniform(20, 60)),
        127 + 100 * np.cos((xx + yy) / rng.uniform(30, 90) + rng.uniform(0, 6)),
        127 + 100 * np.sin(yy / rng.uniform(15, 45) + rng.uniform(0, 6)),
    ], axis=-1)
    noise = rng.integers(-40, 41, size=(h, w, 3)).astype(np.float32)
    return np.clip(np.rint(base + noise), 0, 255).astype(np.uint8)


def synthetic_track(stream: int, t: int = 32, h: int = 720, w: int = 1280, scale: float = 0.5):
    """A t-frame face track: per frame (det box f64[4], lm5 f64[5,2] in frame
    coordinates).  Geometry imitates the shipped fixture
    altfreezing/examples/shining.mp4_32_retina_320.pth (face ~190x220 px,
    ~10 degree roll, slow drift), seed 3000+stream."""
    rng = np.random.default_rng(3000 + int(stream))
    cx, cy = rng.uniform(0.35 * w, 0.65 * w), rng.uniform(0.35 * h, 0.6 * h)
    fw, fh = rng.uniform(150, 230), rng.uniform(190, 270)
    roll = np.deg2rad(rng.uniform(-15, 15))
    vx, vy = rng.uniform(-1.2, 1.2, size=2)
    # canonical 5-point layout in a unit face box (eyes, nose, mouth corners)
    canon = np.array([[0.31, 0.38], [0.69, 0.38], [0.5, 0.57], [0.35, 0.76], [0.65, 0.76]])
    out = []
    for f in range(t):
        jx, jy = rng.normal(0, 0.8, size=2)
        ccx, ccy = cx + vx * f + jx, cy + vy * f + jy
        box = np.array([ccx - fw / 2, ccy - fh / 2, ccx + fw / 2, ccy + fh / 2], np.float64)
        pts = (canon - 0.5) * np.array([fw, fh])
        rot = np.array([[np.cos(roll), -np.sin(roll)], [np.sin(roll), np.cos(roll)]])
        lm5 = pts @ rot.T + np.array([ccx, ccy]) + rng.normal(0, 0.6, size=(5, 2))
        out.append((box, lm5.astype(np.float64)))
    return out
