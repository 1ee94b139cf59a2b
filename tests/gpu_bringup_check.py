"""Bring-up diagnostics on a B200: runs every kernel against torch/oracle references and
prints a table instead of stopping at the first failure.  Usage: python tests/gpu_bringup_check.py [sections]"""
import os
import sys
import time
import traceback

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import afb200  # noqa: E402
from afb200 import synthetic  # noqa: E402
from oracle import crop_oracle, i3d_oracle  # noqa: E402
from tests.helpers import crop_case_inputs  # noqa: E402

dev = torch.device("cuda", 0)


def section(name):
    print("\n=== %s ===" % name, flush=True)


def conv_case(cin, cout, k, stride, pad, B, T, H, W, dtype, impl, relu=True, res=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, H, W, cin, generator=g)
    w = torch.randn(cout, cin, *k, generator=g) * (2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    xd = x.to(dev, dtype)
    xr = xd.float().cpu()                      # reference sees the rounded input
    wr = w.to(dtype).float() if dtype == torch.bfloat16 else w
    ref = F.conv3d(xr.permute(0, 4, 1, 2, 3), wr, b, stride, pad)
    r = None
    if res:
        r = torch.randn(ref.shape[0], *ref.shape[2:], cout, generator=g).to(dev, dtype)
        ref = ref + r.float().cpu().permute(0, 4, 1, 2, 3)
    if relu:
        ref = F.relu(ref)
    ref = ref.permute(0, 2, 3, 4, 1).contiguous()
    t0 = time.time()
    y = afb200.conv_ndhwc(xd, w, b, stride, pad, relu, r, impl=impl)
    torch.cuda.synchronize()
    dt = time.time() - t0
    err = (y.float().cpu() - ref).abs().max().item()
    scale = ref.abs().max().item()
    return err, scale, dt


def run_convs():
    section("conv kernels vs torch fp32 reference")
    cases = [
        # cin, cout, kernel, stride, pad, B, T, H, W
        ("1x1x1 64->256", 64, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 4, 16, 16),
        ("1x1x1 256->64 (M tail)", 256, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 3, 7, 7),
        ("3x1x1 64->64", 64, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 1, 4, 16, 16),
        ("3x1x1 256->128 B2", 256, 128, (3, 1, 1), (1, 1, 1), (1, 0, 0), 2, 4, 8, 8),
        ("1x3x3 64->64", 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), 1, 4, 16, 16),
        ("1x3x3 128->128 B2 7x7", 128, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, 4, 7, 7),
        ("1x3x3 s2 128->128", 128, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1), 1, 4, 16, 16),
        ("1x3x3 s2 64->64 14x14 B2", 64, 64, (1, 3, 3), (1, 2, 2), (0, 1, 1), 2, 2, 14, 14),
        ("1x1x1 s2 256->512", 256, 512, (1, 1, 1), (1, 2, 2), (0, 0, 0), 1, 4, 16, 16),
        ("1x1x1 512->2048 7x7", 512, 2048, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 16, 7, 7),
        ("3x1x1 1024->512 7x7", 1024, 512, (3, 1, 1), (1, 1, 1), (1, 0, 0), 1, 16, 7, 7),
    ]
    for name, cin, cout, k, s, p, B, T, H, W in cases:
        for dtype, impl, tag in ((torch.float32, 1, "simt-f32"), (torch.bfloat16, 1, "simt-bf16"), (torch.bfloat16, 2, "umma-bf16")):
            for res in (False, True):
                try:
                    err, scale, dt = conv_case(cin, cout, k, s, p, B, T, H, W, dtype, impl, relu=True, res=res)
                    tol = 2e-4 if dtype == torch.float32 else 2e-2
                    print("%-28s %-10s res=%d  max|err|=%.3e (scale %.2f) %s  %.1f ms" % (
                        name, tag, res, err, scale, "ok" if err <= tol * max(scale, 1) else "MISMATCH", dt * 1e3), flush=True)
                except Exception as e:
                    print("%-28s %-10s res=%d  ERROR %s" % (name, tag, res, str(e)[:200]), flush=True)
                    if "CUDA" in str(e) or "cuda" in str(e):
                        try:
                            torch.cuda.synchronize()
                        except Exception as e2:
                            print("device is in an error state, stopping conv section:", e2)
                            return False
    return True


def run_crop():
    section("crop kernel vs golden / oracle")
    golden = dict(np.load(os.path.join(ROOT, "tests", "golden", "crop_golden.npz")))
    for name in ("fixture", "synthetic0"):
        lms, imgs, frames, bigs = crop_case_inputs(golden, name, synthetic, afb200.crop)
        lt, wh, diff, tfm, trans = afb200.clip_geometry(bigs, [l[1] for l in lms], 224)
        fr = [torch.from_numpy(f).to(dev) for f in frames]
        out = afb200.crop.crop_u8(fr, bigs, [(tfm, lt, wh)], 32, 224)[0].cpu().numpy()
        want = crop_oracle.crop_align_from_frames(frames, bigs, tfm, lt, wh, 224)
        d = np.abs(out.astype(int) - want.astype(int))
        print("%-10s from-frames kernel: max diff %d, mismatches %d / %d ; golden sub equal: %s" % (
            name, d.max(), (d > 0).sum(), d.size, np.array_equal(out[:, ::4, ::4], golden[name + "_img_sub"])), flush=True)
        t68, img2 = afb200.CropAlignB200(224)(lms, imgs)
        print("%-10s CropAlignB200 wrapper: equal to oracle: %s ; lm68 max diff %.2e" % (
            name, np.array_equal(img2, want), np.abs(t68 - golden[name + "_lm68_t"]).max()), flush=True)


def run_model(precision, n_clips=2):
    section("full model, precision=%s" % precision)
    sd = synthetic.synthetic_state_dict(0)
    golden = dict(np.load(os.path.join(ROOT, "tests", "golden", "model_golden.npz")))
    eng = afb200.Engine(sd, max_batch=4, precision=precision)
    eng.set_option("keep_stages", 1)
    u8 = np.stack([synthetic.synthetic_clip_u8(i) for i in range(n_clips)])
    x = synthetic.normalise_clip(u8).to(dev)
    torch.cuda.synchronize()
    t0 = time.time()
    logits, feats = eng.forward(x, return_features=True)
    torch.cuda.synchronize()
    print("forward %d clips: %.1f ms, launches %d" % (n_clips, (time.time() - t0) * 1e3, eng.launch_count))
    print("logits", logits.view(-1).tolist(), "golden", golden["logits"][:n_clips].ravel().tolist())
    print("max|dlogit| = %.3e ; max|dfeat| = %.3e" % (
        np.abs(logits.cpu().numpy() - golden["logits"][:n_clips]).max(),
        np.abs(feats.cpu().numpy() - golden["features"][:n_clips]).max()))
    _, stages = i3d_oracle.forward(sd, x.cpu(), return_stages=True)
    for si in range(5):
        got = eng.get_stage(si + 1).cpu()
        ref = stages[si]
        print("  stage s%d shape %s  max|err| %.3e  (ref absmax %.3f)  rel-L2 %.3e" % (
            si + 1, tuple(got.shape), (got - ref).abs().max().item(), ref.abs().max().item(),
            ((got - ref).norm() / ref.norm()).item()), flush=True)
    # u8 path must equal the float path on the same clips
    lg2, sc2 = eng.infer_u8(torch.from_numpy(u8).to(dev))
    print("infer_u8 vs forward: max|d| = %.3e" % (lg2.view(-1) - logits.view(-1)).abs().max().item())
    # timing
    for _ in range(2):
        eng.forward(x)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        eng.forward(x)
    torch.cuda.synchronize()
    print("steady: %.2f ms / clip" % ((time.time() - t0) / 3 / n_clips * 1e3), flush=True)
    eng.close()


if __name__ == "__main__":
    want = sys.argv[1:] or ["crop", "convs", "fp32", "bf16"]
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    for sec in want:
        try:
            if sec == "crop":
                run_crop()
            elif sec == "convs":
                if not run_convs():
                    break
            elif sec in ("fp32", "bf16"):
                run_model(sec)
        except Exception:
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e:
                print("device error state:", e)
                break
