"""GPU parity tests (B200): every CUDA kernel, called through the C ABI (ctypes), against the
CPU oracle on the same seeded inputs and against the committed reference-generated goldens.

Tolerances (BASELINE.json north_star): crop <= 1 LSB (we require 0), fp32 path max|dlogit| <= 1e-3,
bf16 path <= 2e-2 with the same decision at logit 0 (after re-centring the head bias).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import afb200
from afb200 import synthetic
from oracle import crop_oracle, i3d_oracle
from tests.helpers import crop_case_inputs, sha256_u8, stage_sample_index

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    assert torch.cuda.get_device_capability(0)[0] == 10, "sm_100a required"
    afb200.lib()                     # fails loudly if libafb200.so is missing
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def clips_u8():
    return np.stack([synthetic.synthetic_clip_u8(i) for i in range(4)])


# ------------------------------------------------------------------ K1 crop / warp / normalise
@pytest.mark.parametrize("name", ["fixture", "synthetic0", "synthetic1", "synthetic2"])
def test_crop_kernel_bit_exact(dev, golden_crop, name):
    lms, imgs, frames, bigs = crop_case_inputs(golden_crop, name, synthetic, afb200.crop)
    lt, wh, diff, tfm, trans = afb200.clip_geometry(bigs, [l[1] for l in lms], 224)
    fr = [torch.from_numpy(f).to(dev) for f in frames]
    out = afb200.crop.crop_u8(fr, bigs, [(tfm, lt, wh)], 32, 224)[0].cpu().numpy()
    assert np.array_equal(sha256_u8(out), golden_crop[name + "_img_sha256"])        # == cv2.warpAffine in the reference
    assert np.array_equal(out[:, ::4, ::4, :], golden_crop[name + "_img_sub"])
    # FasterCropAlignXRay-compatible wrapper on the separate crops
    t68, out2 = afb200.CropAlignB200(224)(lms, imgs)
    assert np.array_equal(out2, out)
    assert np.abs(t68 - golden_crop[name + "_lm68_t"]).max() <= 1e-9


def test_crop_kernel_edge_cases(dev):
    rng = np.random.default_rng(3)
    H, W = 240, 320
    frames = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(4)]
    fr = [torch.from_numpy(f).to(dev) for f in frames]
    cases = [
        # heavy rotation + crop partly outside the canvas; boxes touching the frame border; tiny box
        (np.array([[0.9, -0.6, 40.0], [0.6, 0.9, -70.0]]), [(0, 0, 320, 240)] * 4),
        (np.array([[2.5, 0.0, -300.0], [0.0, 2.5, -200.0]]), [(100, 60, 220, 170), (90, 50, 230, 180), (100, 60, 220, 170), (110, 70, 210, 160)]),
        (np.array([[0.3, 0.0, 10.0], [0.0, 0.3, 10.0]]), [(5, 5, 319, 239)] * 4),
        (np.array([[-1.0, 0.0, 200.0], [0.0, 1.0, 0.0]]), [(150, 100, 153, 102)] * 4),     # reflection, 3x2 box
    ]
    for tfm, boxes in cases:
        bigs = np.array(boxes)
        lt = bigs[:, :2].min(0)
        wh = tuple(int(v) for v in (bigs[:, 2:].max(0) - lt))
        out = afb200.crop.crop_u8(fr, bigs, [(tfm, lt, wh)], 4, 64)[0].cpu().numpy()
        want = crop_oracle.crop_align_from_frames(frames, bigs, tfm, lt, wh, 64)
        assert np.array_equal(out, want)
    # BGR frames: the kernel swaps while reading
    bigs = np.array([(0, 0, 320, 240)] * 4)
    tfm = np.array([[0.8, 0.1, 3.0], [-0.1, 0.8, 5.0]])
    out = afb200.crop.crop_u8(fr, bigs, [(tfm, (0, 0), (320, 240))], 4, 64, bgr=True)[0].cpu().numpy()
    want = crop_oracle.crop_align_from_frames([f[..., ::-1] for f in frames], bigs, tfm, (0, 0), (320, 240), 64)
    assert np.array_equal(out, want)


# ------------------------------------------------------------------ conv kernels, layer level
def _conv_ref(x, w, b, stride, pad, relu, res):
    y = F.conv3d(x.float().cpu().permute(0, 4, 1, 2, 3), w, b, stride, pad)
    if res is not None:
        y = y + res.float().cpu().permute(0, 4, 1, 2, 3)
    return (F.relu(y) if relu else y).permute(0, 2, 3, 4, 1).contiguous()


CONV_CASES = [
    # cin, cout, kernel, stride, pad, B, T, H, W
    (64, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 4, 16, 16),
    (256, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 3, 7, 7),         # M = 147: tail tile
    (64, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 2, 4, 12, 12),
    (128, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, 4, 7, 7),
    (128, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1), 1, 4, 16, 16),
    (256, 512, (1, 1, 1), (1, 2, 2), (0, 0, 0), 1, 4, 14, 14),
    (1024, 512, (3, 1, 1), (1, 1, 1), (1, 0, 0), 1, 16, 7, 7),
    (512, 2048, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 16, 7, 7),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("mode", ["simt_fp32", "simt_bf16", "umma_bf16"])
def test_conv_kernels_vs_torch_fp32(dev, case, mode):
    cin, cout, k, s, p, B, T, H, W = case
    g = torch.Generator().manual_seed(cin * 7 + cout)
    dtype = torch.float32 if mode == "simt_fp32" else torch.bfloat16
    impl = 2 if mode == "umma_bf16" else 1
    x = torch.randn(B, T, H, W, cin, generator=g).to(dev, dtype)
    w = torch.randn(cout, cin, *k, generator=g) * (2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    wr = w.to(dtype).float()
    for res_on in (False, True):
        y0 = _conv_ref(x, wr, b, s, p, True, None)
        res = torch.randn(y0.shape, generator=g).to(dev, dtype) if res_on else None
        want = _conv_ref(x, wr, b, s, p, True, res)
        got = afb200.conv_ndhwc(x, w, b, s, p, True, res, impl=impl).float().cpu()
        # fp32: accumulation-order noise only; bf16: one rounding of the output (2^-8 relative)
        tol = 1e-4 if dtype == torch.float32 else 2.0 ** -8 * max(1.0, want.abs().max().item()) + 1e-3
        assert (got - want).abs().max().item() <= tol


SHORTCUT_CASES = [
    # cin (c conv), cin2 (shortcut), cout, shortcut stride, B, T, Ho, Wo  — first ResBlock of s2 / s3 / s4 / s5
    (64, 64, 256, (1, 1, 1), 1, 3, 12, 12),       # s2.0: stride-1 shortcut (plain 2-D map), M = 432 (tail tile)
    (128, 256, 512, (1, 2, 2), 2, 4, 14, 14),     # s3.0: stride-2 shortcut over a 28x28 input (im2col map)
    (256, 512, 1024, (1, 2, 2), 1, 4, 7, 7),      # s4.0: odd output width, tail tile
    (512, 1024, 2048, (1, 2, 2), 1, 2, 7, 7),     # s5.0
    (64, 128, 64, (1, 2, 2), 1, 2, 5, 9),         # BLOCK_N = 64, odd input extents (H2 = 9, W2 = 17)
]


@pytest.mark.parametrize("case", SHORTCUT_CASES)
def test_fused_projection_shortcut_vs_torch_fp32(dev, case):
    """relu(c(x) + branch1(x2)) as ONE GEMM over K = cin + cin2 (resnet_helper.py:411-423,438-441)."""
    cin, cin2, cout, s2, B, T, Ho, Wo = case
    g = torch.Generator().manual_seed(cin + 3 * cin2 + cout)
    H2 = (Ho - 1) * s2[1] + 1 + (1 if s2[1] > 1 and Ho % 2 == 0 else 0)   # odd and even input extents
    W2 = (Wo - 1) * s2[2] + 1 + (1 if s2[2] > 1 and Wo % 2 == 0 else 0)
    x = torch.randn(B, T, Ho, Wo, cin, generator=g).to(dev, torch.bfloat16)
    x2 = torch.randn(B, T, H2, W2, cin2, generator=g).to(dev, torch.bfloat16)
    w = torch.randn(cout, cin, 1, 1, 1, generator=g) * (1.0 / cin) ** 0.5
    w2 = torch.randn(cout, cin2, 1, 1, 1, generator=g) * (1.0 / cin2) ** 0.5
    b, b2 = torch.randn(cout, generator=g) * 0.1, torch.randn(cout, generator=g) * 0.1
    for relu in (True, False):
        want = _conv_ref(x, w.to(torch.bfloat16).float(), b, (1, 1, 1), (0, 0, 0), False, None) + \
            _conv_ref(x2, w2.to(torch.bfloat16).float(), b2, s2, (0, 0, 0), False, None)
        want = F.relu(want) if relu else want
        got = afb200.conv_shortcut_ndhwc(x, w, b, x2, w2, b2, s2, relu).float().cpu()
        assert got.shape == want.shape
        tol = 2.0 ** -8 * max(1.0, want.abs().max().item()) + 1e-3
        assert (got - want).abs().max().item() <= tol


ROWS_CASES = [
    # cin, cout, kernel, pad, B, T, H, W  (stride 1; row-halo tcgen05 kernel, impl=3)
    (64, 64, (1, 3, 3), (0, 1, 1), 2, 3, 20, 16),      # partial last row tile (20 = 16 + 4)
    (64, 64, (1, 3, 3), (0, 1, 1), 1, 5, 56, 56),      # s2.b geometry
    (64, 64, (5, 4, 1), (2, 0, 0), 1, 6, 35, 24),      # unfolded-stem geometry (Ho = 32)
    (128, 64, (3, 3, 1), (1, 1, 0), 1, 4, 16, 8),      # two channel blocks, temporal + vertical taps
    (64, 64, (1, 2, 3), (0, 0, 1), 3, 1, 17, 8),       # even kernel height, 11 tiles (odd unit count)
]


@pytest.mark.parametrize("case", ROWS_CASES)
def test_rows_kernel_vs_torch_fp32(dev, case):
    cin, cout, k, p, B, T, H, W = case
    g = torch.Generator().manual_seed(cin + 31 * k[0] + 7 * k[1])
    x = torch.randn(B, T, H, W, cin, generator=g).to(dev, torch.bfloat16)
    w = torch.randn(cout, cin, *k, generator=g) * (2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    for relu in (True, False):
        want = _conv_ref(x, w.to(torch.bfloat16).float(), b, (1, 1, 1), p, relu, None)
        got = afb200.conv_ndhwc(x, w, b, (1, 1, 1), p, relu, None, impl=3).float().cpu()
        assert got.shape == want.shape
        tol = 2.0 ** -8 * max(1.0, want.abs().max().item()) + 1e-3
        assert (got - want).abs().max().item() <= tol


@pytest.mark.parametrize("case", [(64, 64, (5, 4, 1), (2, 0, 0), 1, 4, 35, 24),      # stem geometry, Ho=32 Wo=24
                                  (64, 64, (1, 3, 3), (0, 1, 1), 2, 2, 24, 16),      # partial row tile (24 = 16 + 8)
                                  (64, 64, (1, 2, 1), (0, 0, 0), 1, 3, 113, 112)])   # Ho = Wo = 112 like the real stem
def test_rows_kernel_fused_maxpool(dev, case):
    """conv + ReLU + MaxPool3d k[1,3,3] s[1,2,2] p[0,1,1] in one kernel (impl=4) vs torch."""
    cin, cout, k, p, B, T, H, W = case
    g = torch.Generator().manual_seed(17 + H)
    x = torch.randn(B, T, H, W, cin, generator=g).to(dev, torch.bfloat16)
    w = torch.randn(cout, cin, *k, generator=g) * (2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    y = F.relu(F.conv3d(x.float().cpu().permute(0, 4, 1, 2, 3), w.to(torch.bfloat16).float(), b, 1, p))
    y = y.to(torch.bfloat16).float()                     # the kernel pools the bf16-rounded conv output
    want = F.max_pool3d(y, (1, 3, 3), (1, 2, 2), (0, 1, 1)).permute(0, 2, 3, 4, 1).contiguous()
    got = afb200.conv_ndhwc(x, w, b, (1, 1, 1), p, True, None, impl=4).float().cpu()
    assert got.shape == want.shape
    tol = 2.0 ** -8 * max(1.0, want.abs().max().item()) + 1e-3
    assert (got - want).abs().max().item() <= tol


@pytest.mark.parametrize("case", [(64, 256, 1, 4, 8, 8), (128, 64, 2, 2, 16, 12), (64, 256, 1, 32, 56, 56)])
@pytest.mark.parametrize("with_res", [False, True])
def test_pointwise_kernel_fused_temporal_maxpool(dev, case, with_res):
    """1x1x1 conv (+residual) + ReLU + MaxPool3d k=s=[2,1,1] in one kernel (impl=5) vs torch."""
    cin, cout, B, T, H, W = case
    g = torch.Generator().manual_seed(3 + cin + T)
    x = torch.randn(B, T, H, W, cin, generator=g).to(dev, torch.bfloat16)
    w = torch.randn(cout, cin, 1, 1, 1, generator=g) * (2.0 / cin) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    res = torch.randn(B, T, H, W, cout, generator=g).to(dev, torch.bfloat16) if with_res else None
    y = _conv_ref(x, w.to(torch.bfloat16).float(), b, (1, 1, 1), (0, 0, 0), True, res)      # NDHWC fp32
    y = y.to(torch.bfloat16).float().permute(0, 4, 1, 2, 3)
    want = F.max_pool3d(y, (2, 1, 1), (2, 1, 1)).permute(0, 2, 3, 4, 1).contiguous()
    got = afb200.conv_ndhwc(x, w, b, (1, 1, 1), (0, 0, 0), True, res, impl=5).float().cpu()
    assert got.shape == want.shape
    tol = 2.0 ** -8 * max(1.0, want.abs().max().item()) + 1e-3
    assert (got - want).abs().max().item() <= tol


@pytest.mark.parametrize("case", [(64, 1, 4, 8, 8), (256, 2, 8, 12, 12), (128, 1, 4, 56, 56), (256, 1, 32, 20, 7)])
def test_temporal_sweep_kernel_vs_torch_fp32(dev, case):
    """3x1x1 conv, Cout 64 (s2 `a` convs) through the temporal-sweep tcgen05 kernel (impl=6): temporal zero
    padding at both clip ends, several clips, ragged last pixel tile (H*W not a multiple of 128)."""
    cin, B, T, H, W = case
    g = torch.Generator().manual_seed(cin + T + H)
    x = torch.randn(B, T, H, W, cin, generator=g).to(dev, torch.bfloat16)
    w = torch.randn(64, cin, 3, 1, 1, generator=g) * (2.0 / (cin * 3)) ** 0.5
    b = torch.randn(64, generator=g) * 0.1
    want = _conv_ref(x, w.to(torch.bfloat16).float(), b, (1, 1, 1), (1, 0, 0), True, None)
    got = afb200.conv_ndhwc(x, w, b, (1, 1, 1), (1, 0, 0), True, None, impl=6).float().cpu()
    tol = 2.0 ** -8 * max(1.0, want.abs().max().item()) + 1e-3
    assert (got - want).abs().max().item() <= tol


def test_umma_and_simt_bf16_agree_closely(dev):
    """Same bf16 inputs, both fp32-accumulating: results may differ only by accumulation
    order, i.e. by at most one bf16 ulp after the final rounding."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 4, 14, 14, 256, generator=g).to(dev, torch.bfloat16)
    w = torch.randn(256, 256, 1, 3, 3, generator=g) * 0.03
    b = torch.randn(256, generator=g) * 0.1
    a = afb200.conv_ndhwc(x, w, b, (1, 1, 1), (0, 1, 1), True, None, impl=1).float()
    c = afb200.conv_ndhwc(x, w, b, (1, 1, 1), (0, 1, 1), True, None, impl=2).float()
    d = (a - c).abs()
    assert d.max().item() <= 2.0 ** -7 * max(1.0, a.abs().max().item())
    assert (d > 0).float().mean().item() < 0.05


# ------------------------------------------------------------------ whole path
@pytest.fixture(scope="module")
def oracle_out(state_dict, clips_u8):
    x = synthetic.normalise_clip(clips_u8)
    logits, stages = i3d_oracle.forward(state_dict, x, return_stages=True)
    return x, logits, stages


def test_fp32_path_matches_oracle_and_golden(dev, state_dict, clips_u8, oracle_out, golden_model):
    x, o_logits, o_stages = oracle_out
    eng = afb200.Engine(state_dict, max_batch=4, precision="fp32")
    eng.set_option("keep_stages", 1)
    logits, feats = eng.forward(x.to(dev), return_features=True)
    assert (logits.cpu() - o_logits).abs().max().item() <= 1e-3
    assert np.abs(logits.cpu().numpy() - golden_model["logits"]).max() <= 1e-3
    assert np.abs(feats.cpu().numpy() - golden_model["features"]).max() <= 1e-3
    for si, name in enumerate(("s1", "s2", "s3", "s4", "s5")):
        got = eng.get_stage(si + 1).cpu()
        assert tuple(got.shape) == tuple(golden_model[name + "_shape"])
        idx = stage_sample_index(got.numel())
        assert np.abs(got.reshape(-1).numpy()[idx] - golden_model[name + "_samples"]).max() <= 1e-4
        assert (got - o_stages[si]).abs().max().item() <= 1e-4
    eng.close()


def test_bf16_path_within_tolerance_and_same_decision(dev, state_dict, clips_u8, oracle_out, golden_model):
    x, o_logits, o_stages = oracle_out
    # re-centre the head bias so that the four clips straddle the decision threshold
    sd = dict(state_dict)
    med = float(o_logits.median())
    sd["resnet.head.projection.bias"] = state_dict["resnet.head.projection.bias"] - med
    ref = o_logits - med
    eng = afb200.Engine(sd, max_batch=4, precision="bf16")
    eng.set_option("keep_stages", 1)
    logits = eng.forward(x.to(dev)).cpu()
    err = (logits - ref).abs().max().item()
    assert err <= 2e-2, err
    near_tie = ref.abs() < 2e-2
    assert torch.equal((logits > 0)[~near_tie], (ref > 0)[~near_tie])
    assert (ref > 0).any() and (ref < 0).any()
    for si in range(5):
        got = eng.get_stage(si + 1).cpu()
        rel = ((got - o_stages[si]).norm() / o_stages[si].norm()).item()
        assert rel <= 2e-2, (si, rel)
    # the launch counter moves: these were our kernels, not a fallback
    assert eng.launch_count >= 50
    eng.close()
    # production schedule (no kept stages: stem max-pool and temporal max-pool fused into conv epilogues)
    eng2 = afb200.Engine(sd, max_batch=4, precision="bf16")
    logits2 = eng2.forward(x.to(dev)).cpu()
    assert (logits2 - ref).abs().max().item() <= 2e-2
    assert (logits2 - logits).abs().max().item() <= 5e-3
    eng2.close()


def test_input_layouts_and_dtypes_give_same_logits(dev, state_dict, clips_u8):
    eng = afb200.Engine(state_dict, max_batch=2, precision="fp32")
    x = synthetic.normalise_clip(clips_u8[:2]).to(dev)
    base = eng.forward(x)
    # permuted NTHWC view (demo.py:317) and channels_last_3d (TEST2.py:155)
    nthwc = x.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
    assert torch.equal(eng.forward(nthwc), base)
    assert torch.equal(eng.forward(x.contiguous(memory_format=torch.channels_last_3d)), base)
    # u8 entry == float entry when the constants agree
    lg, sc = eng.infer_u8(torch.from_numpy(clips_u8[:2]).to(dev))
    assert (lg.view(-1) - base.view(-1)).abs().max().item() <= 1e-5
    assert torch.allclose(sc, torch.sigmoid(lg))
    # more clips than max_batch are chunked by the host
    x3 = torch.cat([x, x[:1]])
    assert torch.equal(eng.forward(x3)[:2], base)
    with pytest.raises(ValueError):
        eng.forward(x[:, :, :16])
    eng.close()


def test_crop_infer_equals_crop_then_infer(dev, state_dict):
    H, W = 720, 1280
    eng = afb200.Engine(state_dict, max_batch=2, precision="bf16")
    frames, boxes, geoms, fr_all = [], [], [], []
    for c in range(2):
        track = synthetic.synthetic_track(c)
        fr = [torch.from_numpy(synthetic.synthetic_frame_u8(10 * c + f)).to(dev) for f in range(32)]
        bigs = np.stack([afb200.get_crop_box((H, W), b, 0.5) for b, _ in track])
        lm5_rel = [lm - big[:2][None] for (_, lm), big in zip(track, bigs)]
        lt, wh, diff, tfm, trans = afb200.clip_geometry(bigs, lm5_rel, 224)
        frames += fr
        boxes += list(bigs)
        geoms.append((tfm, lt, wh))
    u8 = afb200.crop.crop_u8(frames, boxes, geoms, 32, 224)
    lg_a, sc_a = eng.infer_u8(u8)
    fd, cg = afb200.crop.pack_descriptors(frames, boxes, geoms, dev)
    lg_b, sc_b = eng.crop_infer(fd, cg, 2)
    assert torch.equal(lg_a, lg_b)            # fused path writes the identical normalised clip
    # host-buffer service call (ClassifierSvc.infer_scores boundary)
    scores = eng.infer_scores_u8_host(u8.cpu().numpy())
    assert np.allclose(scores, sc_a.cpu().numpy(), atol=1e-6)
    eng.close()


def test_classifier_plugin_interface_and_feature_hook(dev, state_dict, clips_u8, oracle_out):
    x, o_logits, o_stages = oracle_out
    clf = afb200.Classifier(precision="fp32", max_batch=2).to(dev).eval()
    clf.load_state_dict_tolerant({"state_dict": {"module." + k: v for k, v in state_dict.items()}})
    with torch.no_grad():
        out = clf(x[:2].to(dev))
    assert set(out) == {"final_output"} and tuple(out["final_output"].shape) == (2, 1)
    assert (out["final_output"].cpu() - o_logits[:2]).abs().max().item() <= 1e-3
    # feature.py:106-114: hook the LAST nn.Linear and read its input
    lin = [m for m in clf.modules() if isinstance(m, torch.nn.Linear)][-1]
    grabbed = {}
    h = lin.register_forward_hook(lambda m, i, o: grabbed.update(feat=i[0].detach()))
    with torch.no_grad():
        out2 = clf(x[:2].to(dev))
    h.remove()
    assert grabbed["feat"].reshape(2, -1).shape == (2, 2048)
    assert (grabbed["feat"].reshape(2, -1).cpu() - o_stages[5][:2]).abs().max().item() <= 1e-3
    assert (out2["final_output"] - out["final_output"]).abs().max().item() <= 1e-4
    # reloading other weights re-folds
    sd2 = synthetic.synthetic_state_dict(1)
    clf.load_state_dict_tolerant(sd2)
    with torch.no_grad():
        out3 = clf(x[:1].to(dev))
    assert (out3["final_output"] - out["final_output"][:1]).abs().max().item() > 1e-3


def test_service_infer_scores(dev, state_dict, clips_u8, oracle_out):
    x, o_logits, _ = oracle_out
    svc = afb200.ClassifierSvc(state_dict, precision="bf16", max_batch=2)
    scores = svc.infer_scores(clips_u8[:3])
    assert scores.shape == (3,) and scores.dtype == np.float32
    assert np.abs(scores - torch.sigmoid(o_logits[:3, 0]).numpy()).max() <= 5e-3
    with pytest.raises(ValueError):
        svc.infer_scores(clips_u8[:, :8])


def test_pipelined_host_api_matches_blocking_call(dev, state_dict, clips_u8):
    """af_submit_u8_host / af_wait (two batches in flight) == af_infer_u8_host, batch by batch, incl. a ragged one."""
    svc = afb200.ClassifierSvc(state_dict, device=0, precision="bf16", max_batch=3)
    batches = [clips_u8[:3], clips_u8[1:4], clips_u8[3:4], clips_u8[:2]]
    want = [svc.infer_scores(b).copy() for b in batches]
    got = list(svc.infer_scores_stream(batches))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(g, w)
    # a third submission without collecting the first is refused, not queued over live results
    eng = svc.engine
    pin = torch.from_numpy(np.ascontiguousarray(clips_u8[:1])).pin_memory()
    t0 = eng.submit_u8_host_ptr(pin.data_ptr(), 1)
    t1 = eng.submit_u8_host_ptr(pin.data_ptr(), 1)
    with pytest.raises(afb200.Afb200Error):
        eng.submit_u8_host_ptr(pin.data_ptr(), 1)
    s0, _ = eng.wait(t0, 1)
    s1, _ = eng.wait(t1, 1)
    assert np.array_equal(s0, s1) and np.array_equal(s0, want[0][:1])
    with pytest.raises(afb200.Afb200Error):
        eng.wait(t0, 1)


def test_linearity_property_of_conv_at_full_size(dev):
    """Size-independent property at a full BASELINE layer size (s2 1x3x3, one clip):
    conv(x1 + x2) == conv(x1) + conv(x2) without bias/ReLU, up to bf16 rounding."""
    g = torch.Generator().manual_seed(9)
    x1 = (torch.randn(1, 32, 56, 56, 64, generator=g) * 0.5).to(dev, torch.bfloat16)
    x2 = (torch.randn(1, 32, 56, 56, 64, generator=g) * 0.5).to(dev, torch.bfloat16)
    w = torch.randn(64, 64, 1, 3, 3, generator=g) * 0.05
    z = torch.zeros(64)
    xs = (x1.float() + x2.float()).to(torch.bfloat16)
    xs_err = (xs.float() - (x1.float() + x2.float())).abs().max().item()
    a = afb200.conv_ndhwc(xs, w, z, (1, 1, 1), (0, 1, 1), False, None, impl=2).float()
    b = afb200.conv_ndhwc(x1, w, z, (1, 1, 1), (0, 1, 1), False, None, impl=2).float() + \
        afb200.conv_ndhwc(x2, w, z, (1, 1, 1), (0, 1, 1), False, None, impl=2).float()
    scale = b.abs().max().item()
    assert (a - b).abs().max().item() <= 3 * 2.0 ** -8 * scale + 576 * 0.05 * xs_err


def test_live_ring_scoring_matches_direct_path(dev, state_dict):
    """Streaming layer: windows over a device frame ring give the same scores as crop_u8 -> infer_u8."""
    from afb200 import live
    H, W = 360, 640
    eng = afb200.Engine(state_dict, max_batch=4, precision="bf16")
    ring = live.FrameRing(eng, 48, H, W)
    scorer = live.LiveScorer(live.make_ring_score_fn(eng, ring), clip_size=32, stride=8)
    track = synthetic.synthetic_track(5, t=40, h=H, w=W)
    frames = [torch.from_numpy(synthetic.synthetic_frame_u8(100 + f, H, W)) for f in range(40)]
    got = []
    obs = []
    for f in range(40):
        box, lm = track[f]
        big = afb200.get_crop_box((H, W), box, 0.5)
        slot = ring.put(frames[f])
        obs.append((slot, big, lm - big[:2][None]))
        scorer.observe(0, slot, big, lm - big[:2][None])
        got += scorer.flush()
    assert [round(x) for x in [len(got)]] == [2]                    # windows end at frames 31 and 39
    # direct path for the second window (frames 8..39)
    win = obs[8:40]
    bigs = np.stack([o[1] for o in win])
    lt, wh, diff, tfm, trans = afb200.clip_geometry(bigs, [o[2] for o in win], 224)
    u8 = afb200.crop.crop_u8([frames[f].to(dev) for f in range(8, 40)], bigs, [(tfm, lt, wh)], 32, 224)
    lg, sc = eng.infer_u8(u8)
    assert abs(float(sc[0]) - got[1][1]) <= 1e-6
    eng.close()


def test_odd_batches_and_tail_tiles_are_batch_invariant(dev, state_dict, clips_u8):
    """B=3 (M tails in s5: 3*784 rows is not a multiple of 128; frame-pair tiles; partial chunks) must give
    exactly the per-clip results of B=1 runs: every clip's arithmetic is independent of its batch mates."""
    eng = afb200.Engine(state_dict, max_batch=3, precision="bf16")
    u8 = torch.from_numpy(clips_u8[:3]).to(dev)
    lg3, _ = eng.infer_u8(u8)
    for i in range(3):
        lg1, _ = eng.infer_u8(u8[i:i + 1].contiguous())
        assert abs(float(lg1[0]) - float(lg3[i])) <= 1e-6, i
    # chunk sizes must not change results either
    eng.set_option("chunk_front", 1)
    eng.set_option("chunk_back", 2)
    lg3b, _ = eng.infer_u8(u8)
    assert torch.equal(lg3, lg3b)
    eng.close()


def test_bad_arguments_are_reported_not_crashed(dev, state_dict):
    eng = afb200.Engine(state_dict, max_batch=1, precision="fp32")
    with pytest.raises(afb200.Afb200Error, match="unknown option"):
        eng.set_option("no_such_option", 1)
    with pytest.raises(afb200.Afb200Error, match="not kept"):
        eng.get_stage(2)
    with pytest.raises(afb200.Afb200Error):
        afb200.conv_ndhwc(torch.zeros(1, 2, 8, 8, 6, device=dev), torch.zeros(64, 6, 1, 1, 1), torch.zeros(64),
                          (1, 1, 1), (0, 0, 0), True)          # cin % 4 != 0
    eng.close()


def test_rgb_backbone_adapter_frame_features(dev, state_dict, clips_u8, oracle_out):
    """BASELINE config 5 / SURVEY A13: the AltFreezingRGBEncoder contract — backbone([B,T,3,H,W]) -> [B,T',D], whose
    (masked) temporal mean is the pooled feature.  Oracle: spatial mean of the oracle's last stage."""
    x, o_logits, o_stages = oracle_out
    clf = afb200.Classifier(precision="fp32", max_batch=2).to(dev).eval()
    clf.load_state_dict_tolerant(state_dict)
    backbone = afb200.RGBBackboneB200(clf)
    frames = x[:2].permute(0, 2, 1, 3, 4).contiguous().to(dev)           # [B,T,3,H,W]
    zt = backbone(frames)
    assert tuple(zt.shape) == (2, 16, 2048)
    want = o_stages[4][:2].mean(dim=(3, 4)).permute(0, 2, 1)             # [B,16,2048]
    assert (zt.cpu() - want).abs().max().item() <= 1e-4
    # AltFreezingRGBEncoder.forward: masked temporal mean (dual_rgb.py:37-44)
    pooled = zt.mean(dim=1)
    assert (pooled.cpu() - o_stages[5][:2]).abs().max().item() <= 1e-4
    mask = torch.zeros(2, 16, dtype=torch.bool, device=dev)
    mask[:, 8:] = True
    valid = (~mask).float()
    w = (valid / valid.clamp_min(1e-6).sum(dim=1, keepdim=True)).unsqueeze(-1)
    assert ((zt * w).sum(dim=1).cpu() - want[:, :8].mean(dim=1)).abs().max().item() <= 1e-4


def test_feature_export_infer_clip(dev, state_dict, clips_u8, oracle_out):
    from afb200 import features
    x, o_logits, o_stages = oracle_out
    eng = afb200.Engine(state_dict, max_batch=1, precision="fp32")
    logits, feat, score = features.infer_clip(eng, clips_u8[1])
    assert tuple(logits.shape) == (1, 1) and tuple(feat.shape) == (1, 1, 1, 1, 2048) and tuple(score.shape) == (1,)
    assert abs(float(logits) - float(o_logits[1, 0])) <= 1e-3
    assert (feat.view(-1) - o_stages[5][1]).abs().max().item() <= 1e-3
    assert abs(float(score) - float(torch.sigmoid(o_logits[1, 0]))) <= 1e-4
    eng.close()
