"""Penultimate-feature export: the `.npz` per-clip record that
altfreezing/feature.py writes (`process_video.flush`, feature.py:190-214) and that dualrun's RGB branch consumes.

  AFModel.infer_clip(aligned u8 [T,H,W,3]) -> (logits [1,1], feat [1,1,1,1,2048], score [1])   feature.py:116-150
The reference grabs `feat` with a forward hook on the last nn.Linear; here the engine returns the pooled features
directly (`af_infer_u8` features output), shaped like the hook's input ([B,1,1,1,2048] after the NTHWC permute).
"""
import os

import numpy as np
import torch


def infer_clip(engine, aligned_thwc_u8: np.ndarray):
    """-> (logits cpu [1,1], feat cpu [1,1,1,1,2048], score cpu [1])"""
    x = torch.from_numpy(np.ascontiguousarray(aligned_thwc_u8, dtype=np.uint8)).unsqueeze(0).to(engine.device)
    logits, scores, feats = engine.infer_u8(x, return_features=True)
    return logits.view(1, 1).cpu(), feats.view(1, 1, 1, 1, -1).cpu(), scores.view(1).cpu()


def save_clip_npz(path, logits, feat, score, y, tid, clip_idx, video_rel, save_fp16=True):
    """Same keys and dtypes as feature.py:198-207."""
    dt = np.float16 if save_fp16 else np.float32
    np.savez_compressed(path, feat=feat.numpy().astype(dt), logits=logits.numpy().astype(dt),
                        score=float(score.squeeze().item()), y=np.int64(y), tid=np.int64(tid),
                        clip_idx=np.int64(clip_idx), video_rel=video_rel)
    return path


def clip_npz_name(out_dir, vname, tid, clip_idx):
    return os.path.join(out_dir, "%s_tid%d_c%05d.npz" % (vname, tid, clip_idx))
