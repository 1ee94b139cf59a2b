"""Video-level decision and the per-video / summary CSV wire format.

  video_decision           <- VideoRunner.run, altfreezing/TEST2.py:696-744 (per-track pooling via live.pool_track,
                              stability penalty, threshold, low-quality q75/q90 rule, video_score = max raw track score)
  PER_VIDEO_HEADER, per_video_row, write_per_video
                           <- altfreezing/TEST2.py:1070-1076,1095-1105,1118-1119 — the file altfreezing/ds.py:35-58
                              (`load_per_video`) and dualrun/rgb/engine_rgb.py:245-261 read back by column NAME
  SUMMARY_HEADER, summary_row, write_summary
                           <- altfreezing/TEST2.py:1121-1149
Everything here is host-side bookkeeping around scores the engine produced; formatting (fixed decimals, "nan"
spelling, human-readable model size) is kept character for character because downstream scripts parse it.
"""
import csv
import math
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from .live import pool_track, score_with_stability

PER_VIDEO_HEADER = [
    "video_path", "dataset", "subset", "gt_label", "pred_label", "correct",
    "video_score", "threshold",
    "frames_processed", "elapsed_s", "fps", "latency_ms_clip_mean",
    "num_tracks", "id_switch_rate_per_1k_frames",
    "gpu_mem_alloc_peak_mb", "gpu_mem_reserved_peak_mb", "cpu_mem_peak_mb", "model_size",
]
SUMMARY_HEADER = [
    "videos", "accuracy", "auc_roc", "pr_auc", "f1",
    "tp", "tn", "fp", "fn", "confusion_matrix", "mean_fps", "mean_latency_ms_clip",
    "model_size",
]


def human_bytes(n) -> str:
    """altfreezing/TEST2.py:119-123."""
    n = float(n)
    for unit in ["B", "KB", "MB", "GB", "TB"]:
        if n < 1024.0:
            return f"{n:.1f}{unit}"
        n /= 1024.0
    return f"{n:.1f}PB"


def video_decision(track_clip_scores: Dict[int, Sequence[float]], threshold: float = 0.0, pool_method: str = "median",
                   topk_ratio: float = 0.2, percentile_p: float = 80.0, trim_ratio: float = 0.2, min_clips: int = 1,
                   disable_penalty: bool = False, low_quality: bool = False, qa_q75_thr: float = 1.0,
                   qa_q90_thr: float = 1.0) -> dict:
    """Per-track clip scores -> the video verdict of VideoRunner.run (TEST2.py:696-744): each track with at least
    `min_clips` clips is pooled (`pool_track`), optionally penalised for instability, and called fake above `threshold`;
    on low-quality videos a track is also fake when its q75 / q90 reach the QA thresholds; the video is fake if any track
    is, and `video_score` (what AUC is computed on) is the maximum RAW pooled track score."""
    raw_scores, per_person = {}, {}
    for tid, scores in track_clip_scores.items():
        if len(scores) < min_clips:
            continue
        raw = pool_track(scores, method=pool_method, topk_ratio=topk_ratio, percentile_p=percentile_p, trim_ratio=trim_ratio)
        pen = raw if disable_penalty else score_with_stability(scores, raw)
        raw_scores[tid], per_person[tid] = float(raw), float(pen)
    quants = {}
    for tid, ss in track_clip_scores.items():
        s = np.asarray(ss, float)
        if s.size:
            q = np.percentile(s, [10, 25, 50, 75, 90])
            quants[tid] = {"q10": q[0], "q25": q[1], "q50": q[2], "q75": q[3], "q90": q[4]}
    labels = {}
    for tid in per_person:
        std = int(per_person[tid] > threshold)
        q = quants.get(tid)
        qa = int(bool(low_quality and q and (q["q75"] >= qa_q75_thr or q["q90"] >= qa_q90_thr)))
        labels[tid] = int(std or qa)
    return {"pred_label": int(any(v == 1 for v in labels.values())),
            "video_score": float(max(raw_scores.values())) if raw_scores else 0.0,
            "per_track_raw": raw_scores, "per_track_penalised": per_person, "per_track_label": labels,
            "per_track_quantiles": quants}


def _f(v, fmt) -> str:
    return "nan" if (isinstance(v, float) and math.isnan(v)) else format(v, fmt)


def per_video_row(video_path: str, dataset: str, subset: str, gt_label: int, res: dict, threshold,
                  model_size_bytes: int = 0) -> list:
    """One row of the per-video CSV (TEST2.py:1095-1105).  `res` carries pred_label, video_score, frames_processed,
    elapsed_s, fps, latency_ms_clip_mean, num_tracks, id_switch_rate_per_1k_frames and the three memory peaks
    (float('nan') where unknown)."""
    pred = int(res["pred_label"])
    score = float(res["video_score"])
    return [
        video_path, dataset, subset, gt_label, pred, int(pred == gt_label),
        f"{score:.6f}", threshold,
        res["frames_processed"], f"{res['elapsed_s']:.3f}", f"{res['fps']:.3f}",
        _f(float(res["latency_ms_clip_mean"]), ".3f"),
        res["num_tracks"], f"{res['id_switch_rate_per_1k_frames']:.3f}",
        _f(float(res.get("gpu_mem_alloc_peak_mb", float("nan"))), ".1f"),
        _f(float(res.get("gpu_mem_reserved_peak_mb", float("nan"))), ".1f"),
        _f(float(res.get("cpu_mem_peak_mb", float("nan"))), ".1f"),
        human_bytes(model_size_bytes),
    ]


def write_per_video(path: str, rows: Iterable[list]) -> None:
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(PER_VIDEO_HEADER)
        w.writerows(rows)


def summary_row(rows: List[list], model_size_bytes: int = 0) -> list:
    """The one-line summary CSV (TEST2.py:1121-1145): accuracy, ROC-AUC (nan with a single class), PR-AUC and F1 via
    scikit-learn when it is importable (nan otherwise, as in the reference), the confusion counts and mean fps/latency."""
    y_true = [int(r[3]) for r in rows]
    y_pred = [int(r[4]) for r in rows]
    y_score = [float(r[6]) for r in rows]
    tp = sum(1 for g, p in zip(y_true, y_pred) if g == 1 and p == 1)
    tn = sum(1 for g, p in zip(y_true, y_pred) if g == 0 and p == 0)
    fp = sum(1 for g, p in zip(y_true, y_pred) if g == 0 and p == 1)
    fn = sum(1 for g, p in zip(y_true, y_pred) if g == 1 and p == 0)
    acc = f1 = auc = ap = float("nan")
    cm = [[0, 0], [0, 0]]
    try:
        from sklearn.metrics import accuracy_score, average_precision_score, confusion_matrix, f1_score, roc_auc_score
        if y_true:
            acc = accuracy_score(y_true, y_pred)
            f1 = f1_score(y_true, y_pred)
            auc = roc_auc_score(y_true, y_score) if len(set(y_true)) > 1 else float("nan")
            ap = average_precision_score(y_true, y_score)
            cm = confusion_matrix(y_true, y_pred).tolist()
    except ImportError:
        pass
    mean_fps = float(np.nanmean([float(r[10]) for r in rows])) if rows else float("nan")
    lats = [float(r[11]) if r[11] != "nan" else np.nan for r in rows]
    mean_lat = float(np.nanmean(lats)) if rows and not all(np.isnan(lats)) else float("nan")
    return [len(rows), _f(float(acc), ".6f"), _f(float(auc), ".6f"), _f(float(ap), ".6f"), _f(float(f1), ".6f"),
            tp, tn, fp, fn, cm, _f(mean_fps, ".3f"), _f(mean_lat, ".3f"), human_bytes(model_size_bytes)]


def write_summary(path: str, rows: List[list], model_size_bytes: int = 0) -> list:
    row = summary_row(rows, model_size_bytes)
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(SUMMARY_HEADER)
        w.writerow(row)
    return row


def read_scores(path: str) -> Dict[str, float]:
    """The consumer side as dualrun/rgb/engine_rgb.py:245-261 does it: video stem and `technique/stem` -> video_score."""
    import os
    out: Dict[str, float] = {}
    with open(path, newline="") as f:
        for r in csv.DictReader(f):
            vp = r["video_path"]
            stem = os.path.splitext(os.path.basename(vp))[0]
            tech = os.path.basename(os.path.dirname(vp))
            out[stem] = out[f"{tech}/{stem}"] = float(r["video_score"])
    return out
