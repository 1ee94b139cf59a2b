"""CPU tests: the oracle against the reference-generated golden vectors, the host logic, and
the C-ABI surface (load + exported symbols, no compute)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import afb200
from afb200 import synthetic
from oracle import crop_oracle, i3d_oracle
from tests.helpers import crop_case_inputs, sha256_u8, stage_sample_index

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_model_oracle_matches_reference_golden(golden_model, state_dict):
    idx = [0, 3]
    u8 = np.stack([synthetic.synthetic_clip_u8(i) for i in idx])
    logits, stages = i3d_oracle.forward(state_dict, synthetic.normalise_clip(u8), return_stages=True)
    ref = golden_model["logits"][idx]
    assert np.abs(logits.numpy() - ref).max() <= 2e-5
    assert np.abs(stages[5].numpy() - golden_model["features"][idx]).max() <= 2e-5
    # stage samples: golden was sampled over the 4-clip tensor; compare where the sample falls in our clips
    for si, name in enumerate(("s1", "s2", "s3", "s4", "s5")):
        shape = golden_model[name + "_shape"]
        per_clip = int(np.prod(shape[1:]))
        gi = stage_sample_index(int(np.prod(shape)))
        clip_of = gi // per_clip
        ours = stages[si].numpy().reshape(len(idx), -1)
        for j, c in enumerate(idx):
            sel = clip_of == c
            got = ours[j][gi[sel] - c * per_clip]
            want = golden_model[name + "_samples"][sel]
            assert sel.sum() > 100
            assert np.abs(got - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), name


def test_synthetic_weights_are_reference_schema(state_dict):
    keys = afb200.network.reference_key_set()
    assert len(keys) == 320
    assert set(keys) == set(state_dict)
    for k, shp in keys.items():
        assert tuple(state_dict[k].shape) == shp, k
    net = afb200.I3D8x8Params()
    net.load_state_dict(state_dict, strict=True)
    assert sum(p.numel() for p in net.parameters()) == 27225921


def test_arch_table_counts():
    specs = afb200.arch.all_conv_specs()
    assert len(specs) == 53
    assert afb200.arch.macs_per_clip() == 113627365376
    kts = [b.a.kernel[0] for b in afb200.arch.block_specs()]
    assert kts == [3, 3, 3, 3, 1, 3, 1, 3, 1, 3, 1, 3, 1, 1, 3, 1]


@pytest.mark.parametrize("name", ["fixture", "synthetic0", "synthetic1", "synthetic2"])
def test_crop_oracle_matches_reference_golden(golden_crop, name):
    lms, imgs, frames, bigs = crop_case_inputs(golden_crop, name, synthetic, crop_oracle)
    lm68_t, out = crop_oracle.crop_align(lms, imgs, 224)
    assert np.array_equal(sha256_u8(out), golden_crop[name + "_img_sha256"])
    assert np.array_equal(out[:, ::4, ::4, :], golden_crop[name + "_img_sub"])
    assert np.abs(lm68_t - golden_crop[name + "_lm68_t"]).max() <= 1e-9
    # zero-copy formulation: gather from the full frames, masked to each frame's own box
    lt, wh, diff, tfm, trans = crop_oracle.clip_geometry(bigs, [l[1] for l in lms], 224)
    assert np.array_equal(bigs, golden_crop[name + "_big_boxes"])
    assert np.abs(tfm - golden_crop[name + "_tfm"]).max() <= 1e-12
    out2 = crop_oracle.crop_align_from_frames(frames, bigs, tfm, lt, wh, 224)
    assert np.array_equal(out, out2)


def test_warp_affine_oracle_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for it in range(6):
        h, w = rng.integers(230, 330, 2)
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ang, sc = rng.uniform(-0.6, 0.6), rng.uniform(0.5, 1.8)
        M = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-80, 40)],
                      [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-150, 40)]])
        assert np.array_equal(cv2.warpAffine(src, M, (224, 224)), crop_oracle.warp_affine_u8(src, M, 224))


@pytest.mark.parametrize("name", ["fixture", "synthetic0", "synthetic1", "synthetic2"])
def test_host_geometry_matches_reference_golden(golden_crop, name):
    """Product host logic (afb200.crop) — get_crop_box, union box, similarity fit."""
    lms, imgs, frames, bigs = crop_case_inputs(golden_crop, name, synthetic, afb200.crop)
    assert np.array_equal(bigs, golden_crop[name + "_big_boxes"])
    lt, wh, diff, tfm, trans = afb200.clip_geometry(bigs, [l[1] for l in lms], 224)
    assert np.array_equal(np.asarray(lt), golden_crop[name + "_left_top"])
    assert np.abs(tfm - golden_crop[name + "_tfm"]).max() <= 1e-12
    assert np.abs(trans - golden_crop[name + "_trans"]).max() <= 1e-12


def test_get_crop_box_edges():
    # clipping at the frame border and rint-half-even behaviour, vs the oracle restatement
    rng = np.random.default_rng(5)
    for _ in range(200):
        h, w = rng.integers(100, 1100, 2)
        x1, y1 = rng.uniform(-20, w), rng.uniform(-20, h)
        box = np.array([x1, y1, x1 + rng.uniform(1, 400), y1 + rng.uniform(1, 400)])
        box = np.round(box * 2) / 2          # provoke .5 ties
        sc = float(rng.choice([0.0, 0.3, 0.5]))
        assert np.array_equal(afb200.get_crop_box((h, w), box, sc), crop_oracle.get_crop_box((h, w), box, sc))


def test_similarity_degenerate_points_raise():
    pts = np.zeros((2, 5, 2))
    with pytest.raises(Exception):
        afb200.estimate_clip_transform(pts, afb200.crop.STD_POINTS_256)


def test_bn_fold_matches_conv_bn(state_dict):
    spec = afb200.arch.block_specs()[4].b          # a stride-2 1x3x3
    w, b = afb200.fold_conv_bn(state_dict, spec)
    x = torch.randn(1, spec.cin, 2, 10, 10)
    want = i3d_oracle._bn(state_dict, spec.bn, i3d_oracle._conv(state_dict, spec.name, x, spec.stride, spec.pad))
    got = torch.nn.functional.conv3d(x, torch.from_numpy(w), torch.from_numpy(b), spec.stride, spec.pad)
    assert (got - want).abs().max() <= 1e-5


def test_checkpoint_unwrapping_and_tolerant_load(tmp_path, state_dict):
    sd = {("module." + k): v for k, v in state_dict.items()}
    sd["module.extra.key"] = torch.zeros(3)
    sd["module.resnet.head.projection.bias"] = torch.zeros(7)          # wrong shape -> skipped
    path = tmp_path / "ckpt.pth"
    torch.save({"classifier_state_dict": sd}, path)
    clf = afb200.Classifier()
    ok, epoch = clf.load(str(path), epoch=3)
    assert ok and epoch == 3
    own = clf.network.state_dict()
    k = "resnet.s3.pathway0_res1.branch2.b.weight"
    assert torch.equal(own[k], state_dict[k])
    assert own["resnet.head.projection.bias"].shape == (1,)
    assert clf.load(str(tmp_path / "missing.pth")) == (False, -1)
    assert "network.resnet.head.projection.weight" in clf.state_dict()
    assert not any(k.startswith("_warped_network") for k in clf.state_dict())
    # parameters() is redirected to the network (model/_base.py:174-175)
    assert sum(p.numel() for p in clf.parameters()) == 27225921


def test_cpu_input_fails_loudly(state_dict):
    clf = afb200.Classifier().eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        clf(torch.zeros(1, 3, 32, 224, 224))


def test_c_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "afb200.h")).read()
    declared = set(re.findall(r"\b(af_[a-z0-9_]+)\s*\(", header))
    declared -= {"af_status", "af_handle"}
    assert declared == set(afb200._lib.EXPORTS), declared ^ set(afb200._lib.EXPORTS)
    assert os.path.exists(afb200.LIB_PATH), "libafb200.so not built: run python __graft_entry__.py"
    L = ctypes.CDLL(afb200.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    L.af_version.restype = ctypes.c_int32
    assert L.af_version() == 102
    assert afb200.lib().af_last_error() is not None


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every struct in include/afb200.h as gcc sees them == the ctypes mirrors in _lib.py."""
    import shutil
    import subprocess
    L = afb200._lib
    structs = {"af_conv_desc": L.AfConvDesc, "af_block_desc": L.AfBlockDesc, "af_tt_layer": L.AfTTLayer,
               "af_tt_head": L.AfTTHead, "af_weights": L.AfWeights, "af_frame_desc": L.AfFrameDesc,
               "af_clip_geom": L.AfClipGeom}
    assert ctypes.sizeof(L.AfConvDesc) == 16 + 11 * 4 + 4 and ctypes.sizeof(L.AfBlockDesc) == 24
    assert ctypes.sizeof(L.AfFrameDesc) == 40 and ctypes.sizeof(L.AfClipGeom) == 64
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "afb200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_folded_weights_struct(state_dict):
    fw = afb200.FoldedWeights(state_dict)
    assert fw.struct.n_convs == 53 and fw.struct.n_blocks == 16 and fw.struct.stem == 0
    pools = [fw.blocks[i].temporal_pool_before for i in range(16)]
    assert pools == [0, 0, 0, 1] + [0] * 12
    assert [fw.blocks[i].branch1 >= 0 for i in range(16)] == [True, False, False, True, False, False, False,
                                                              True, False, False, False, False, False, True, False, False]


def test_vectorised_geometry_equals_the_per_clip_routines():
    """get_crop_boxes / clip_geometry_batch (host side of a 32-clip step) == get_crop_box / clip_geometry clip by clip,
    incl. a mirrored track (the reference's reflected-candidate branch, warp_for_xray.py:395-425)."""
    H, W = 720, 1280
    boxes, lms = [], []
    for c in range(6):
        track = afb200.synthetic.synthetic_track(40 + c)
        det = np.stack([b for b, _ in track])
        if c == 2:
            det[:, [0, 2]] -= 900.0                       # runs off the left frame edge: clipping
        bigs = afb200.get_crop_boxes((H, W), det, 0.5)
        assert np.array_equal(bigs, np.stack([afb200.get_crop_box((H, W), b, 0.5) for b in det]))
        lm = np.stack([l - big[:2][None] for (_, l), big in zip(track, bigs)])
        if c == 4:                                        # mirror the face inside its box
            lm = lm.copy()
            lm[..., 0] = (bigs[:, 2] - bigs[:, 0])[:, None] - lm[..., 0]
        boxes.append(bigs)
        lms.append(lm)
    geoms = afb200.clip_geometry_batch(np.stack(boxes), np.stack(lms), 224)
    mirrored = 0
    for c in range(6):
        lt, wh, diff, tfm, trans = afb200.clip_geometry(boxes[c], lms[c], 224)
        assert np.array_equal(lt, geoms[c][1]) and tuple(wh) == tuple(geoms[c][2])
        assert np.abs(tfm - geoms[c][0]).max() <= 1e-12
        mirrored += int(np.linalg.det(tfm[:, :2]) < 0)
    assert mirrored == 1
