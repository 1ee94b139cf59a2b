"""Streaming layer around the hot path: per-track sliding 32-frame
windows over device-resident frame rings, micro-batched clip scoring, and the reference's
score pooling / decisions.

  TrackWindows.push / due        <- RealtimeAF.step window logic, test/af_realtime.py:450-479
                                     (buffer of clip_size frames per track, emit every `stride` frames)
  LiveScorer.flush               <- RealtimeAF._flush_and_infer, test/af_realtime.py:318-360
                                     (crop-align per clip, one batched infer_scores, median-of-5 hysteresis 0.75/0.65)
  pool_track                     <- VideoRunner._pool_track, altfreezing/TEST2.py:636-683 (8 pooling methods)
  score_with_stability           <- altfreezing/TEST2.py:627-634
  decide_meeting_fake            <- test/app_realtime.py:75-92 (80th percentile >= 0.362 after >= 128 frames)

Unlike the reference, frames are uploaded once into a per-stream ring on the GPU and a clip is just
32 (slot, box) descriptors + one 2x3 transform: the crop kernel gathers from the ring, so overlapping
windows share their frames and no aligned u8 clip ever crosses PCIe.
"""
import collections
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .crop import clip_geometry, clip_geometry_batch, get_crop_box


# ------------------------------------------------------------------ score pooling / decisions (host)
def pool_track(scores, method="median", topk_ratio=0.2, percentile_p=80.0, trim_ratio=0.2) -> float:
    s = np.asarray(scores, float)
    if s.size == 0:
        return 0.0

    def logit_median(v):
        se = np.clip(v, 1e-6, 1 - 1e-6)
        med = np.median(np.log(se / (1 - se)))
        return float(1 / (1 + np.exp(-med)))

    if method == "mean":
        return float(np.mean(s))
    if method == "median":
        return float(np.median(s))
    if method == "logit_median":
        return logit_median(s)
    if method in ("topk", "topk_median"):
        k = max(1, int(np.ceil(topk_ratio * s.size)))
        top = np.sort(s)[-k:]
        return float(np.mean(top) if method == "topk" else np.median(top))
    if method == "percentile":
        return float(np.percentile(s, float(np.clip(percentile_p, 0.0, 100.0))))
    if method == "trimmed_mean":
        t = float(np.clip(trim_ratio, 0.0, 0.49))
        ss = np.sort(s)
        a = int(ss.size * t)
        b = max(a + 1, ss.size - a)
        return float(np.mean(ss[a:b]))
    if method == "adaptive":
        iqr = np.percentile(s, 75) - np.percentile(s, 25)
        if iqr < 0.15:
            return float(np.percentile(s, float(np.clip(percentile_p, 0.0, 100.0))))
        return logit_median(s)
    return float(np.median(s))


def score_with_stability(scores, base: float) -> float:
    s = np.asarray(scores, float)
    if s.size == 0:
        return 0.0
    iqr = np.percentile(s, 85) - np.percentile(s, 25)
    if iqr > 0.25 and np.median(s) < 0.85:
        return base * (0.85 ** (iqr / 0.25))
    return base


def decide_meeting_fake(running_scores: Dict[int, List[float]], frames_per_tid: Dict[int, int],
                        threshold: float = 0.362, min_frames: int = 128, percentile_p: float = 80.0):
    """-> (ready, is_fake)."""
    any_ready = False
    for tid, sc in running_scores.items():
        if int(frames_per_tid.get(tid, 0)) >= min_frames and sc:
            any_ready = True
            if float(np.percentile(np.asarray(sc, float), percentile_p)) >= threshold:
                return True, True
    return (True, False) if any_ready else (False, False)


class Hysteresis:
    """Median of the last 5 clip scores with enter/leave thresholds 0.75 / 0.65 per track."""

    def __init__(self, t_high=0.75, t_low=0.65, history=5):
        self.t_high, self.t_low = t_high, t_low
        self.hist = collections.defaultdict(lambda: collections.deque(maxlen=history))
        self.fake: Dict[int, bool] = {}

    def update(self, tid, score: float) -> bool:
        self.hist[tid].append(float(score))
        sm = float(np.median(self.hist[tid]))
        st = self.fake.get(tid, False)
        if not st and sm >= self.t_high:
            st = True
        elif st and sm < self.t_low:
            st = False
        self.fake[tid] = st
        return st


# ------------------------------------------------------------------ sliding windows
class TrackWindows:
    """Per-track buffers of the last `clip_size` observations; a window is due when the buffer is full
    and at least `stride` frames were pushed since the track's last emission."""

    def __init__(self, clip_size=32, stride=8):
        self.clip_size, self.stride = clip_size, stride
        self.buf = collections.defaultdict(lambda: collections.deque(maxlen=clip_size))
        self.since_emit = collections.Counter()
        self.frames_per_tid = collections.Counter()

    def push(self, tid, obs) -> bool:
        """obs: (ring_slot, big_box[4], lm5_rel[5,2]).  Returns True when a window is due."""
        self.buf[tid].append(obs)
        self.frames_per_tid[tid] += 1
        self.since_emit[tid] += 1
        full = len(self.buf[tid]) == self.clip_size
        # the first full window is emitted at once (reference: buffer == clip_size), later ones every `stride`
        if full and (self.since_emit[tid] >= self.stride or self.frames_per_tid[tid] == self.clip_size):
            self.since_emit[tid] = 0
            return True
        return False

    def window(self, tid) -> list:
        return list(self.buf[tid])

    def drop(self, tid):
        self.buf.pop(tid, None)
        self.since_emit.pop(tid, None)


class LiveScorer:
    """Collects due windows, scores them in one batched call and applies the per-track decisions.
    `score_fn(clips) -> scores` takes a list of clips, each a list of (ring_slot, big_box, lm5_rel)."""

    def __init__(self, score_fn: Callable[[List[list]], Sequence[float]], clip_size=32, stride=8,
                 score_is_real=False, max_batch=32):
        self.windows = TrackWindows(clip_size, stride)
        self.score_fn, self.score_is_real, self.max_batch = score_fn, score_is_real, max_batch
        self.pending: List[Tuple[int, list]] = []
        self.running_scores = collections.defaultdict(list)
        self.hyst = Hysteresis()

    def observe(self, tid, ring_slot, big_box, lm5_rel):
        if self.windows.push(tid, (ring_slot, big_box, lm5_rel)):
            self.pending.append((tid, self.windows.window(tid)))

    def flush(self):
        """-> [(tid, score, is_fake)] for every pending window (empty list if none)."""
        out = []
        while self.pending:
            batch, self.pending = self.pending[: self.max_batch], self.pending[self.max_batch:]
            scores = np.asarray(self.score_fn([w for _, w in batch]), dtype=float)
            if self.score_is_real:
                scores = 1.0 - scores
            for (tid, _), s in zip(batch, scores):
                self.running_scores[tid].append(float(s))
                out.append((tid, float(s), self.hyst.update(tid, float(s))))
        return out

    def meeting_decision(self, threshold=0.362, min_frames=128, percentile_p=80.0):
        return decide_meeting_fake(self.running_scores, self.windows.frames_per_tid, threshold, min_frames, percentile_p)


# ------------------------------------------------------------------ device frame ring + clip scoring
class FrameRing:
    """A ring of decoded frames in device memory (one per stream): frames are uploaded once and shared by
    all overlapping windows that reference them."""

    def __init__(self, engine, slots: int, height: int, width: int):
        import torch
        self.engine, self.slots, self.h, self.w = engine, slots, height, width
        self.buf = torch.empty((slots, height, width, 3), dtype=torch.uint8, device=engine.device)
        self.next = 0

    def put(self, frame_u8) -> int:
        """frame_u8: [H,W,3] u8 torch tensor (pinned host or device) -> ring slot."""
        slot = self.next % self.slots
        self.buf[slot].copy_(frame_u8, non_blocking=True)
        self.next += 1
        return slot


    def put_rows(self, slots, host_ptrs, row0, row1, stream=None):
        """Batched feed (af_ring_put_rows): rows [row0[i], row1[i]) of the pinned host frame at address host_ptrs[i]
        -> ring slot slots[i], all queued on `stream` (a torch.cuda.Stream; default: the current one)."""
        import ctypes as C
        import torch
        from ._lib import check, lib
        n = len(slots)
        sl = np.ascontiguousarray(slots, np.int32)
        hp = np.ascontiguousarray(host_ptrs, np.uint64)
        r0 = np.ascontiguousarray(row0, np.int32)
        r1 = np.ascontiguousarray(row1, np.int32)
        st = stream if stream is not None else torch.cuda.current_stream(self.buf.device)
        with torch.cuda.device(self.buf.device):
            check(lib().af_ring_put_rows(C.c_void_p(self.buf.data_ptr()), self.buf.stride(0), self.buf.stride(1), n,
                                         C.c_void_p(sl.ctypes.data), C.c_void_p(hp.ctypes.data), C.c_void_p(r0.ctypes.data),
                                         C.c_void_p(r1.ctypes.data), C.c_void_p(st.cuda_stream)), "af_ring_put_rows")


    def put_boxes(self, slots, host_ptrs, boxes_xyxy, stream=None):
        """Batched feed of the face boxes only (af_ring_put_boxes): pixels boxes_xyxy[i] = (x0, y0, x1, y1) of the
        pinned host frame at address host_ptrs[i] -> the same pixels of ring slot slots[i], one strided copy each."""
        import ctypes as C
        import torch
        from ._lib import check, lib
        sl = np.ascontiguousarray(slots, np.int32)
        hp = np.ascontiguousarray(host_ptrs, np.uint64)
        bx = np.ascontiguousarray(boxes_xyxy, np.int32).reshape(-1, 4)
        assert len(sl) == len(hp) == len(bx)
        st = stream if stream is not None else torch.cuda.current_stream(self.buf.device)
        with torch.cuda.device(self.buf.device):
            check(lib().af_ring_put_boxes(C.c_void_p(self.buf.data_ptr()), self.buf.stride(0), self.buf.stride(1), len(sl),
                                          C.c_void_p(sl.ctypes.data), C.c_void_p(hp.ctypes.data), C.c_void_p(bx.ctypes.data),
                                          C.c_void_p(st.cuda_stream)), "af_ring_put_boxes")


def ring_descriptors(ring: FrameRing, clips: List[list], size: int = 224):
    """Descriptors (device arrays) for windows whose frames all live in `ring`."""
    from .crop import pack_descriptors_ring
    slots = [o[0] for win in clips for o in win]
    boxes = [np.stack([np.asarray(o[1]) for o in win]) for win in clips]
    if len({len(win) for win in clips}) == 1:        # equal-length windows: one vectorised geometry pass
        geoms = clip_geometry_batch(np.stack(boxes), np.stack([np.stack([o[2] for o in win]) for win in clips]), size)
    else:
        geoms = []
        for win, bigs in zip(clips, boxes):
            lt, wh, diff, tfm, trans = clip_geometry(bigs, [o[2] for o in win], size)
            geoms.append((tfm, lt, wh))
    buf = ring.buf
    return pack_descriptors_ring(buf.data_ptr(), buf.stride(0), buf.stride(1), ring.h, ring.w, slots,
                                 np.concatenate(boxes), geoms, ring.engine.device)


def make_ring_score_fn(engine, ring: FrameRing, size: int = 224, bgr: bool = False):
    """score_fn for LiveScorer: build descriptors for the windows and run the fused crop+trunk call."""

    def score(clips: List[list]):
        fd, cg = ring_descriptors(ring, clips, size)
        logits, scores = engine.crop_infer(fd, cg, len(clips), bgr=bgr)
        return scores.cpu().numpy()
    return score
