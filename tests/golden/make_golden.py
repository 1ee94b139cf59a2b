"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, via oracle/ref_loader.py) in the build container, and checks the
oracle restatements against it while doing so.  Run:  python tests/golden/make_golden.py

The reference holds no golden vectors or tests for this path, so these
reference-generated fixtures are what pins the oracle.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import afb200  # noqa: E402
from afb200 import synthetic  # noqa: E402
from oracle import crop_oracle, i3d_oracle, ref_loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
STAGE_SAMPLES = 4096


def stage_sample_index(numel, n=STAGE_SAMPLES):
    return (np.arange(n, dtype=np.int64) * 2654435761 + 12345) % numel


def golden_model():
    torch.manual_seed(0)
    sd = synthetic.synthetic_state_dict(0)
    clf = ref_loader.reference_classifier()
    missing = clf.network.load_state_dict(sd, strict=True)
    print("loaded synthetic weights into the reference network:", missing)
    n_clips = 4
    u8 = np.stack([synthetic.synthetic_clip_u8(i) for i in range(n_clips)])
    x = synthetic.normalise_clip(u8)
    feats = []
    hook = clf.network.resnet.head.projection.register_forward_hook(lambda m, i, o: feats.append(i[0].detach()))
    stage_out = {}
    hooks = [hook]
    for name in ("s1", "s2", "s3", "s4", "s5"):
        mod = getattr(clf.network.resnet, name)
        hooks.append(mod.register_forward_hook(
            lambda m, i, o, name=name: stage_out.setdefault(name, []).append(o[0].detach())))
    logits = []
    with torch.no_grad():
        for i in range(n_clips):
            logits.append(clf(x[i:i + 1])["final_output"])
    for h in hooks:
        h.remove()
    logits = torch.cat(logits).numpy()
    feats = torch.cat([f.reshape(1, -1) for f in feats]).numpy()
    print("reference logits", logits.ravel())

    o_logits, o_stages = i3d_oracle.forward(sd, x, return_stages=True)
    d = np.abs(o_logits.numpy() - logits).max()
    print("oracle vs reference  max|dlogit| = %.3e" % d)
    assert d <= 2e-5, d
    out = {"logits": logits.astype(np.float32), "features": feats.astype(np.float32),
           "weights_seed": np.int64(0), "clip_indices": np.arange(n_clips)}
    for si, name in enumerate(("s1", "s2", "s3", "s4", "s5")):
        ref = torch.cat(stage_out[name]).numpy()
        orc = o_stages[si].numpy()
        rel = np.abs(ref - orc).max() / max(np.abs(ref).max(), 1e-9)
        print("  stage %s shape %s absmax %.3f mean|x| %.4f  oracle rel err %.2e" % (
            name, ref.shape, np.abs(ref).max(), np.abs(ref).mean(), rel))
        assert rel <= 1e-5
        idx = stage_sample_index(ref.size)
        out[name + "_samples"] = ref.ravel()[idx].astype(np.float32)
        out[name + "_shape"] = np.array(ref.shape)
        out[name + "_absmean"] = np.float64(np.abs(ref.astype(np.float64)).mean())
    d = np.abs(o_stages[5].numpy() - feats).max()
    assert d <= 1e-5, d
    np.savez_compressed(os.path.join(OUT, "model_golden.npz"), **out)


def fixture_track():
    """boxes/landmarks of the reference's shipped fixture
    altfreezing/examples/shining.mp4_32_retina_320.pth (32 frames, first face)."""
    p = os.path.join(ref_loader.ALTFREEZING_DIR, "examples", "shining.mp4_32_retina_320.pth")
    detect_res, all_lm68 = torch.load(p, weights_only=False)
    boxes = np.stack([np.asarray(detect_res[i][0][0], np.float64) for i in range(32)])
    lm5 = np.stack([np.asarray(detect_res[i][0][1], np.float64) for i in range(32)])
    lm68 = np.stack([np.asarray(all_lm68[i][0], np.float32) for i in range(32)])
    return boxes, lm5, lm68


def golden_crop():
    import cv2
    ref_crop = ref_loader.reference_crop_align(224)
    ref_gcb = ref_loader.reference_get_crop_box()
    out = {}
    boxes, lm5, lm68 = fixture_track()
    out["fixture_boxes"], out["fixture_lm5"], out["fixture_lm68"] = boxes, lm5, lm68
    H, W = 720, 1280
    cases = {"fixture": (boxes, lm5, lm68)}
    for s in range(3):
        tr = synthetic.synthetic_track(s)
        b = np.stack([t[0] for t in tr])
        l5 = np.stack([t[1] for t in tr])
        # fake lm68: the 5 points tiled with offsets (only transformed, never used for pixels)
        l68 = np.concatenate([l5 + k for k in range(14)], axis=1)[:, :68].astype(np.float32)
        cases["synthetic%d" % s] = (b, l5, l68)
    for name, (b, l5, l68) in cases.items():
        frames = [synthetic.synthetic_frame_u8(f) for f in range(32)]
        lms, imgs, bigs = [], [], []
        for i in range(32):
            big = ref_gcb((H, W), b[i], 0.5)
            assert (big == crop_oracle.get_crop_box((H, W), b[i], 0.5)).all()
            tl = big[:2][None, :]
            lms.append((b[i], l5[i] - tl, l68[i] - tl, big))
            imgs.append(frames[i][big[1]:big[3], big[0]:big[2]])
            bigs.append(big)
        r68, rimg = ref_crop(lms, imgs)
        o68, oimg = crop_oracle.crop_align(lms, imgs, 224)
        assert np.array_equal(rimg, oimg), name
        assert np.abs(r68 - o68).max() == 0.0, name
        lt, wh, diff, tfm, trans = crop_oracle.clip_geometry(bigs, [l[1] for l in lms], 224)
        fimg = crop_oracle.crop_align_from_frames(frames, bigs, tfm, lt, wh, 224)
        assert np.array_equal(rimg, fimg), name
        print("crop case %-10s bit-exact vs reference (canvas and from-frames); tfm=%s" % (name, np.round(tfm, 4).tolist()))
        out[name + "_big_boxes"] = np.stack(bigs).astype(np.int64)
        out[name + "_tfm"] = tfm
        out[name + "_trans"] = trans
        out[name + "_left_top"] = np.asarray(lt, np.int64)
        out[name + "_lm68_t"] = r68
        out[name + "_img_sub"] = rimg[:, ::4, ::4, :].copy()
        out[name + "_img_sha256"] = np.frombuffer(hashlib.sha256(rimg.tobytes()).digest(), np.uint8)
        if name != "fixture":
            out[name + "_boxes"], out[name + "_lm5"], out[name + "_lm68"] = b, l5, l68
    # warpAffine emulation vs cv2 on random transforms incl. pure sub-pixel translations
    rng = np.random.default_rng(7)
    bad = 0
    for it in range(24):
        h, w = rng.integers(230, 400, 2)
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if it < 8:
            M = np.array([[1, 0, -rng.integers(0, 32) / 32. - 3], [0, 1, -rng.integers(0, 32) / 32. - 2]])
        else:
            ang, sc = rng.uniform(-0.6, 0.6), rng.uniform(0.5, 1.8)
            M = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-80, 40)],
                          [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-150, 40)]])
        bad += int((cv2.warpAffine(src, M, (224, 224)) != crop_oracle.warp_affine_u8(src, M, 224)).sum())
    print("warpAffine emulation vs cv2 %s: %d mismatching samples" % (cv2.__version__, bad))
    assert bad == 0
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(OUT, "crop_golden.npz"), **out)


if __name__ == "__main__":
    assert ref_loader.reference_available(), "needs /root/reference"
    golden_crop()
    golden_model()
    print("golden fixtures written to", OUT)
