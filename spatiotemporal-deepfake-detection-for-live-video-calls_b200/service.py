"""Batching services on top of the engine, shaped like the streaming callers' singletons.

  ClassifierSvc.infer_scores  <- altfreezing/TEST2.py:138-204, test/af_realtime.py:64-96
  CropAlignSvc.__call__       <- altfreezing/TEST2.py:207-212
"""
from typing import Optional

import numpy as np
import torch

from .crop import CropAlignB200
from .engine import Engine, mean_std_255


class ClassifierSvc:
    """`infer_scores(u8[B,T,H,W,3]) -> np.float32[B]` (sigmoid of the logit).

    The reference converts to fp32 on the device, permutes to NCTHW, normalises and runs the
    model; here the u8 batch is uploaded once and K-packed, normalised and classified by the
    C library (af_infer_u8_host)."""

    def __init__(self, state_dict, device: int = 0, precision: str = "bf16", max_batch: int = 32,
                 clip_size: int = 32, imsize: int = 224):
        self.engine = Engine(state_dict, device=device, max_batch=max_batch, precision=precision,
                             clip_t=clip_size, clip_s=imsize)
        # TEST2.py:147-148 builds mean/std as float32(mean)*255 on the device
        self.engine.mean255, self.engine.std255 = mean_std_255("svc")
        self.clip_size, self.imsize = clip_size, imsize
        self._last_scores: Optional[np.ndarray] = None
        # The reference keeps `_last_logits` only for 2-class heads and sets it to None for the 1-logit head
        # (TEST2.py:184-199); TEST2.py:1107-1110 softmaxes it when it is not None, so it must stay None here.
        self._last_logits = None
        self.last_logits_1d: Optional[np.ndarray] = None      # the raw logits of the last call, float32 [B]

    def infer_scores(self, aligned_batch_bthwc) -> np.ndarray:
        arr = np.asarray(aligned_batch_bthwc)
        if arr.ndim != 5 or arr.shape[1:] != (self.clip_size, self.imsize, self.imsize, 3):
            raise ValueError("infer_scores expects u8 [B,%d,%d,%d,3], got %s" %
                             (self.clip_size, self.imsize, self.imsize, arr.shape))
        if arr.dtype != np.uint8:
            # the reference casts whatever it gets to fp32; aligned crops are always u8
            arr = np.clip(np.rint(arr), 0, 255).astype(np.uint8)
        scores, logits = self.engine.infer_scores_u8_host(arr, return_logits=True)
        self._last_scores, self.last_logits_1d = scores.copy(), logits
        return scores

    def infer_scores_stream(self, batches):
        """Generator form for callers that score batch after batch (TEST2.py:393-439's flush loop over a whole
        video, batch_eval-style offline scoring): yields the same float32[B] per input batch, but the upload of batch
        i+1 overlaps the compute of batch i (af_submit_u8_host / af_wait).  Batches of at most max_batch clips."""
        pending = None                      # (ticket, batch size, pinned tensor kept alive)
        for arr in batches:
            arr = np.ascontiguousarray(arr, dtype=np.uint8)
            if arr.ndim != 5 or arr.shape[1:] != (self.clip_size, self.imsize, self.imsize, 3) or \
                    arr.shape[0] > self.engine.max_batch:
                raise ValueError("infer_scores_stream expects u8 [B<=%d,%d,%d,%d,3], got %s" %
                                 (self.engine.max_batch, self.clip_size, self.imsize, self.imsize, arr.shape))
            pin = torch.from_numpy(arr).pin_memory()
            ticket = self.engine.submit_u8_host_ptr(pin.data_ptr(), arr.shape[0])
            if pending is not None:
                yield self.engine.wait(pending[0], pending[1])[0]
            pending = (ticket, arr.shape[0], pin)
        if pending is not None:
            yield self.engine.wait(pending[0], pending[1])[0]


class CropAlignSvc:
    def __init__(self, imsize: int = 224, device: int = 0):
        self.fn = CropAlignB200(imsize, device=device)

    def __call__(self, infos, imgs):
        """-> (lm68_T, aligned u8 [T,S,S,3]), the 2-tuple both callers unpack as `_, aligned = svc(infos, imgs)`
        (TEST2.py:207-212,401; test/af_realtime.py:104,325)."""
        return self.fn(infos, imgs)
