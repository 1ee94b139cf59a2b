"""Shared helpers for the parity tests."""
import hashlib

import numpy as np


def stage_sample_index(numel, n=4096):
    # must match tests/golden/make_golden.py
    return (np.arange(n, dtype=np.int64) * 2654435761 + 12345) % numel


def sha256_u8(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def crop_case_inputs(golden, name, synthetic, crop_mod):
    """Rebuild the (landmarks, crops, frames, big boxes) of a golden crop case."""
    H, W = 720, 1280
    if name == "fixture":
        boxes, lm5, lm68 = golden["fixture_boxes"], golden["fixture_lm5"], golden["fixture_lm68"]
    else:
        boxes, lm5, lm68 = golden[name + "_boxes"], golden[name + "_lm5"], golden[name + "_lm68"]
    frames = [synthetic.synthetic_frame_u8(f) for f in range(32)]
    lms, imgs, bigs = [], [], []
    for i in range(32):
        big = crop_mod.get_crop_box((H, W), boxes[i], 0.5)
        tl = big[:2][None, :]
        lms.append((boxes[i], lm5[i] - tl, lm68[i] - tl, big))
        imgs.append(frames[i][big[1]:big[3], big[0]:big[2]])
        bigs.append(big)
    return lms, imgs, frames, np.stack(bigs)
