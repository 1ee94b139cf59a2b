"""Host logic of the streaming layer (no GPU): windows, stride, pooling, hysteresis, meeting decision."""
import numpy as np

from afb200 import live


def test_windows_emit_when_full_then_every_stride():
    tw = live.TrackWindows(clip_size=32, stride=8)
    due = [f for f in range(80) if tw.push(7, (f % 40, (0, 0, 10, 10), np.zeros((5, 2))))]
    assert due == [31, 39, 47, 55, 63, 71, 79]
    win = tw.window(7)
    assert len(win) == 32 and win[-1][0] == 79 % 40 and win[0][0] == 48 % 40
    assert tw.frames_per_tid[7] == 80


def test_scorer_batches_and_applies_hysteresis():
    calls = []

    def score_fn(clips):
        calls.append(len(clips))
        return [0.9] * len(clips)

    sc = live.LiveScorer(score_fn, clip_size=4, stride=2, max_batch=3)
    for f in range(6):
        for tid in range(5):
            sc.observe(tid, f, (0, 0, 4, 4), np.zeros((5, 2)))
        res = sc.flush()
        if f == 3:
            assert len(res) == 5 and calls[-2:] == [3, 2]           # 5 windows, max_batch 3
            assert all(fake for _, _, fake in res)                  # 0.9 >= 0.75 on the first score
    assert sc.meeting_decision(threshold=0.362, min_frames=6) == (True, True)
    assert sc.meeting_decision(threshold=0.95, min_frames=6) == (True, False)
    assert sc.meeting_decision(min_frames=128) == (False, False)


def test_hysteresis_thresholds():
    h = live.Hysteresis()
    seq = [0.5, 0.8, 0.9, 0.9, 0.7, 0.6, 0.6, 0.6, 0.6]
    states = [h.update(1, s) for s in seq]
    # medians: .5 .65 .8 .85 .8 .8 .7 .6 .6 -> enter at >= .75, leave at < .65
    assert states == [False, False, True, True, True, True, True, False, False]


def test_pool_methods_match_numpy_definitions():
    s = np.array([0.1, 0.2, 0.9, 0.95, 0.3, 0.7, 0.8, 0.05, 0.6, 0.65])
    assert live.pool_track(s, "mean") == float(np.mean(s))
    assert live.pool_track(s, "median") == float(np.median(s))
    assert live.pool_track(s, "topk") == float(np.mean(np.sort(s)[-2:]))
    assert live.pool_track(s, "topk_median") == float(np.median(np.sort(s)[-2:]))
    assert live.pool_track(s, "percentile", percentile_p=80) == float(np.percentile(s, 80))
    assert live.pool_track(s, "trimmed_mean") == float(np.mean(np.sort(s)[2:8]))
    lm = live.pool_track(s, "logit_median")
    assert abs(lm - float(np.median(s))) < 0.05
    assert live.pool_track(s, "adaptive") == lm                      # wide IQR -> logit median
    tight = np.array([0.5, 0.52, 0.51, 0.5, 0.53])
    assert live.pool_track(tight, "adaptive") == float(np.percentile(tight, 80))
    assert live.pool_track([], "mean") == 0.0
    assert live.pool_track(s, "unknown") == float(np.median(s))
    # stability penalty only for unstable series with a moderate median
    assert live.score_with_stability(tight, 0.7) == 0.7
    assert live.score_with_stability(s, 0.7) < 0.7


def test_feature_npz_record_schema(tmp_path):
    """Wire format of altfreezing/feature.py:198-207 (consumed by dualrun): keys, dtypes, shapes."""
    import torch
    from afb200 import features
    path = features.clip_npz_name(str(tmp_path), "vid", 3, 7)
    assert path.endswith("vid_tid3_c00007.npz")
    features.save_clip_npz(path, torch.tensor([[0.25]]), torch.arange(2048.).view(1, 1, 1, 1, -1), torch.tensor([0.56]),
                           y=1, tid=3, clip_idx=7, video_rel="fake/vid.mp4")
    z = np.load(path)
    assert set(z.files) == {"feat", "logits", "score", "y", "tid", "clip_idx", "video_rel"}
    assert z["feat"].dtype == np.float16 and z["feat"].shape == (1, 1, 1, 1, 2048)
    assert z["logits"].dtype == np.float16 and z["logits"].shape == (1, 1)
    assert abs(float(z["score"]) - 0.56) < 1e-6 and int(z["y"]) == 1 and int(z["tid"]) == 3 and int(z["clip_idx"]) == 7
    assert str(z["video_rel"]) == "fake/vid.mp4"
