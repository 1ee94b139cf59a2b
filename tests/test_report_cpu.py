"""CPU tests of the video-level decision and the per-video / summary CSV wire format.
Where the reference tree is present (build container) the format is pinned against the reference itself: the header
lists are parsed out of altfreezing/TEST2.py and the file we write is read back by the reference's own consumer,
altfreezing/ds.py:load_per_video."""
import ast
import csv
import importlib.util
import math
import os
import re

import numpy as np
import pytest

import afb200
from afb200 import report

REF = "/root/reference/altfreezing"
have_ref = os.path.isfile(os.path.join(REF, "TEST2.py"))


def _res(score, pred, **kw):
    d = dict(pred_label=pred, video_score=score, frames_processed=300, elapsed_s=12.3456, fps=24.3, latency_ms_clip_mean=7.25,
             num_tracks=2, id_switch_rate_per_1k_frames=0.0, gpu_mem_alloc_peak_mb=1234.56, gpu_mem_reserved_peak_mb=float("nan"),
             cpu_mem_peak_mb=812.0)
    d.update(kw)
    return d


def test_video_decision_rules():
    tracks = {1: [0.1, 0.2, 0.15, 0.12], 2: [0.7, 0.9, 0.8, 0.85, 0.95], 3: [0.99]}
    d = report.video_decision(tracks, threshold=0.5, min_clips=2)
    assert set(d["per_track_raw"]) == {1, 2}                        # track 3 has too few clips
    assert d["per_track_label"] == {1: 0, 2: 1} and d["pred_label"] == 1
    assert d["video_score"] == pytest.approx(np.median(tracks[2]))  # max RAW pooled score
    # unstable series are penalised before thresholding (score_with_stability), the raw score is not
    wild = {7: [0.05, 0.95, 0.1, 0.9, 0.2, 0.8, 0.6, 0.61]}
    d = report.video_decision(wild, threshold=0.5, pool_method="mean")
    raw = float(np.mean(wild[7]))
    assert d["video_score"] == pytest.approx(raw) and d["per_track_penalised"][7] < raw
    assert report.video_decision(wild, threshold=0.5, pool_method="mean", disable_penalty=True)["per_track_penalised"][7] == pytest.approx(raw)
    # low-quality rule: q75 / q90 can flip a track that the threshold alone would pass as real
    calm = {4: [0.30, 0.32, 0.31, 0.45, 0.46]}
    assert report.video_decision(calm, threshold=0.5)["pred_label"] == 0
    assert report.video_decision(calm, threshold=0.5, low_quality=True, qa_q75_thr=0.4, qa_q90_thr=0.9)["pred_label"] == 1
    assert report.video_decision({}, threshold=0.5) == {"pred_label": 0, "video_score": 0.0, "per_track_raw": {},
                                                        "per_track_penalised": {}, "per_track_label": {},
                                                        "per_track_quantiles": {}}


def test_per_video_row_formatting_and_roundtrip(tmp_path):
    rows = [report.per_video_row("/data/ffpp/Deepfakes/000_003.mp4", "ffpp", "test", 1, _res(0.812345678, 1), 0.362, 104_200_000),
            report.per_video_row("/data/ffpp/original/000.mp4", "ffpp", "test", 0, _res(0.1, 1, latency_ms_clip_mean=float("nan")), 0.362, 104_200_000)]
    assert rows[0][6] == "0.812346" and rows[0][9] == "12.346" and rows[0][10] == "24.300" and rows[0][11] == "7.250"
    assert rows[0][14] == "1234.6" and rows[0][15] == "nan" and rows[0][17] == "99.4MB" and rows[0][5] == 1
    assert rows[1][5] == 0 and rows[1][11] == "nan"
    p = tmp_path / "per_video.csv"
    report.write_per_video(str(p), rows)
    with open(p, newline="") as f:
        rd = list(csv.DictReader(f))
    assert list(rd[0].keys()) == report.PER_VIDEO_HEADER and float(rd[0]["video_score"]) == pytest.approx(0.812346)
    sc = report.read_scores(str(p))
    assert sc["000_003"] == pytest.approx(0.812346) and sc["Deepfakes/000_003"] == sc["000_003"] and "original/000" in sc
    s = report.write_summary(str(tmp_path / "summary.csv"), rows, 104_200_000)
    assert s[0] == 2 and (s[5], s[6], s[7], s[8]) == (1, 0, 1, 0) and s[10] == "24.300" and s[11] == "7.250"
    assert s[1] == "0.500000" and s[2] == "1.000000"                       # accuracy, AUC with sklearn present
    with open(tmp_path / "summary.csv", newline="") as f:
        assert next(csv.reader(f)) == report.SUMMARY_HEADER
    assert report.human_bytes(0) == "0.0B" and report.human_bytes(1536) == "1.5KB"


@pytest.mark.skipif(not have_ref, reason="reference tree not present")
def test_wire_format_against_the_reference(tmp_path):
    src = open(os.path.join(REF, "TEST2.py"), encoding="utf-8").read()
    m = re.search(r"\n    header = (\[.*?\])\n", src, re.S)
    assert ast.literal_eval(m.group(1)) == report.PER_VIDEO_HEADER
    m = re.search(r"summary_header = (\[.*?\])\n", src, re.S)
    assert ast.literal_eval(m.group(1)) == report.SUMMARY_HEADER
    # the reference's own consumer reads our file back
    rows = [report.per_video_row("/d/ffpp/Face2Face/%03d.mp4" % i, "ffpp", "test", i % 2, _res(0.1 + 0.2 * i, int(i > 1)), 0.362)
            for i in range(4)]
    p = tmp_path / "per_video.csv"
    report.write_per_video(str(p), rows)
    spec = importlib.util.spec_from_file_location("_ref_ds", os.path.join(REF, "ds.py"))
    ds = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(ds)
    except ImportError as e:                      # its plotting / stats dependencies may be absent
        pytest.skip("reference ds.py does not import here: %s" % e)
    y_true, y_score, fps, lat, gpu_a, gpu_r, cpu_m = ds.load_per_video(str(p))
    assert y_true.tolist() == [0, 1, 0, 1] and np.allclose(y_score, [0.1, 0.3, 0.5, 0.7])
    assert np.allclose(fps, 24.3) and np.allclose(lat, 7.25) and np.isnan(gpu_r).all() and np.allclose(cpu_m, 812.0)
    # the nested pooling / penalty helpers of VideoRunner.run, lifted out of the reference source, agree with ours
    m = re.search(r"\n        def score_with_stability\(scores, base\):\n(.*?)\n\n        def _pool_track", src, re.S)
    ns = {"np": np}
    exec("def score_with_stability(scores, base):\n" + "\n".join(l[8:] for l in m.group(1).split("\n")), ns)
    rng = np.random.default_rng(0)
    for _ in range(20):
        s = rng.uniform(0, 1, size=int(rng.integers(1, 40))).tolist()
        assert afb200.live.score_with_stability(s, 0.7) == pytest.approx(ns["score_with_stability"](s, 0.7))
    # _pool_track (TEST2.py:636-683): all eight methods + the fallback, lifted the same way
    m = re.search(r"\n        def _pool_track\((.*?)\):\n(.*?)\n            # fallback\n            return float\(np\.median\(s\)\)\n", src, re.S)
    assert m, "reference _pool_track not found"
    body = m.group(2) + "\n            # fallback\n            return float(np.median(s))"
    exec("def _pool_track(" + m.group(1) + "):\n" + "\n".join(l[8:] for l in body.split("\n")), ns)
    methods = ["mean", "median", "logit_median", "topk", "topk_median", "percentile", "trimmed_mean", "adaptive", "nonsense"]
    for trial in range(30):
        n = int(rng.integers(1, 60))
        s = (rng.uniform(0.4, 0.55, size=n) if trial % 3 == 0 else rng.uniform(0, 1, size=n)).tolist()   # tight and wide IQRs
        for meth in methods:
            kw = dict(topk_ratio=float(rng.choice([0.1, 0.2, 0.5])), percentile_p=float(rng.choice([50.0, 80.0, 95.0, 120.0])),
                      trim_ratio=float(rng.choice([0.0, 0.2, 0.6])))
            want = ns["_pool_track"](s, meth, **kw)
            got = afb200.live.pool_track(s, meth, **kw)
            assert got == pytest.approx(want, rel=1e-12, abs=1e-15), (meth, kw, n)
    assert afb200.live.pool_track([], "mean") == ns["_pool_track"]([], "mean") == 0.0
