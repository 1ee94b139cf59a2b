#!/usr/bin/env python
"""Benchmark of the clip-classification hot path (BASELINE.json metric: clips/s for 32x224^2
face-crop clips; p50 batch-1 latency).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one pass of the hot path over one batch of synthetic input:
  workload = BASELINE.json configs[1]: bf16, batch 32 clips per GPU, crop/warp/normalise kernel
  (from decoded 720p frames + face boxes already resident in HBM) followed by the I3D ResNet-50
  trunk and the head.  `value` is device-timed (CUDA events, max over ranks) with inputs in
  HBM; `e2e` is the same metric through the host-buffer C-ABI calls at the CROP boundary (pinned
  decoded frames -> af_ring_put_boxes -> af_crop_infer -> scores on the host, H2D/D2H inside the
  timed region: the work the reference arm does); `e2e_aligned` is the ClassifierSvc.infer_scores
  boundary (pinned u8 aligned clips in, scores out).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

_emit = print          # replaced by _QuietStdout.emit when run as a script
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "clips_per_s_32x224x224"
UNIT = "clips/s"
FLOPS_PER_CLIP = 2 * 113627365376          # 53 convs, SURVEY.md §8d / afb200.arch.macs_per_clip()
K1_BYTES_PER_CLIP = 18.7e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--workload", default="clips", choices=["clips", "offline", "live", "rgb"],
                    help="clips: BASELINE configs[1] (the headline; also runs short forms of the others unless --no-extra); "
                         "offline / live / rgb: BASELINE configs 3 / 4 / 5 on their own")
    ap.add_argument("--offline-clips", type=int, default=4096)
    ap.add_argument("--live-streams", type=int, default=64)
    ap.add_argument("--live-seconds", type=float, default=4.0)
    ap.add_argument("--rgb-batch", type=int, default=64)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short offline / live / rgb workloads")
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "tf32"])
    ap.add_argument("--variant", default="i3d", choices=["i3d", "ftcn_tt"],
                    help="i3d: the AltFreezing I3D (BASELINE.json's metric); ftcn_tt: the reference's second classifier "
                         "plugin (SURVEY 8f row 4) through the same pipeline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-kernel-events", action="store_true", help="do not bracket each conv launch with CUDA events")
    ap.add_argument("--chunk-front", type=int, default=0)
    ap.add_argument("--chunk-back", type=int, default=0)
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_from=None, t_to=None):
        """Summarise the samples that arrived in [t_from, t_to] (wall clock); all samples if no window given."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, r in self.rows:
            if t_from is not None and not (t_from <= ts <= t_to):
                continue
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- CPU legs (oracle port)
def cpu_reference_step(sd, clip_inputs, n_clips, variant="i3d"):
    """The reference's CPU path for n_clips clips, batch 1 each (demo.py:309-328):
    crop/align (cv2.warpAffine, the reference's own dependency; numpy emulation if cv2 is
    absent) -> normalise -> fp32 forward (oracle port of the reference network)."""
    import numpy as np
    import torch
    from afb200 import synthetic
    from oracle import crop_oracle, ftcn_oracle, i3d_oracle
    net_forward = ftcn_oracle.forward if variant == "ftcn_tt" else i3d_oracle.forward
    try:
        import cv2
    except Exception:
        cv2 = None
    out = []
    for i in range(n_clips):
        frames, bigs, tfm, lt, wh = clip_inputs[i % len(clip_inputs)]
        imgs = []
        for f, bb in zip(frames, bigs):
            canvas = np.zeros((wh[1], wh[0], 3), np.uint8)
            x, y = bb[0] - lt[0], bb[1] - lt[1]
            crop = f[bb[1]:bb[3], bb[0]:bb[2]]
            canvas[y:y + crop.shape[0], x:x + crop.shape[1]] = crop
            imgs.append(cv2.warpAffine(canvas, tfm, (224, 224)) if cv2 is not None
                        else crop_oracle.warp_affine_u8(canvas, tfm, 224))
        x = synthetic.normalise_clip(np.stack(imgs))
        out.append(float(torch.sigmoid(net_forward(sd, x))[0, 0]))
    return out


def make_cpu_clip_inputs(n):
    import numpy as np
    import afb200
    from afb200 import synthetic
    H, W = 720, 1280
    res = []
    for s in range(n):
        track = synthetic.synthetic_track(s)
        frames = [synthetic.synthetic_frame_u8(f) for f in range(32)]
        bigs = np.stack([afb200.get_crop_box((H, W), b, 0.5) for b, _ in track])
        lm5_rel = [lm - big[:2][None] for (_, lm), big in zip(track, bigs)]
        lt, wh, diff, tfm, trans = afb200.clip_geometry(bigs, lm5_rel, 224)
        res.append((frames, bigs, tfm, lt, wh))
    return res


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the
    reference tree itself is Python and is not present on the GPU box), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from afb200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synthetic.synthetic_state_dict(0, args.variant)
    inputs = make_cpu_clip_inputs(1)
    clips_per_step = 1
    for _ in range(args.warmup):
        cpu_reference_step(sd, inputs, clips_per_step, args.variant)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(sd, inputs, clips_per_step, args.variant)
    dt = time.perf_counter() - t0
    value = clips_per_step * args.steps / dt
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "%s clip classification: crop/warp/normalise + trunk, "
                                   "32x224x224 clips (reference CPU path, batch 1 per step: demo.py loop)" % _net_name(args.variant),
                       "clips_per_step": clips_per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d steps x %d clip (cv2.warpAffine crop + fp32 torch forward of the oracle port)" % (args.steps, clips_per_step)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(json.dumps(line))


# --------------------------------------------------------------------------- our arm
def build_gpu_inputs(dev, batch, rank):
    """Frames + descriptors resident in HBM for `batch` clips: a ring of 256+ distinct 720p
    frames (sliding 32-frame windows, stride 8, as in the live-call path) and per-clip geometry
    from seeded synthetic tracks.  Also returns the (frames, boxes, geoms) lists for the parity leg."""
    import numpy as np
    import torch
    import afb200
    from afb200 import synthetic
    H, W = 720, 1280
    n_frames = 8 * batch + 32
    g = torch.Generator(device=dev).manual_seed(2000 + rank)
    pool = torch.randint(0, 256, (n_frames, H, W, 3), dtype=torch.uint8, device=dev, generator=g)
    frames, boxes, geoms = [], [], []
    for c in range(batch):
        track = synthetic.synthetic_track(1000 * rank + c)
        bigs = np.stack([afb200.get_crop_box((H, W), b, 0.5) for b, _ in track])
        lm5_rel = [lm - big[:2][None] for (_, lm), big in zip(track, bigs)]
        lt, wh, diff, tfm, trans = afb200.clip_geometry(bigs, lm5_rel, 224)
        for t in range(32):
            frames.append(pool[8 * c + t])
            boxes.append(bigs[t])
        geoms.append((tfm, lt, wh))
    fd, cg = afb200.crop.pack_descriptors(frames, boxes, geoms, dev)
    src_bytes = sum(int((b[2] - b[0]) * (b[3] - b[1]) * 3) for b in boxes)
    return pool, fd, cg, src_bytes, (frames, boxes, geoms)


def _traffic_for(variant, B, precision):
    """DRAM bytes per step of the conv launches from the committed ncu capture of THIS variant (None if absent)."""
    names = ["r02_traffic.json", "r01_traffic.json"] if variant == "i3d" else ["r02_traffic_%s.json" % variant]
    for nm in names:
        tp = os.path.join(ROOT, "profiles", nm)
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("batch") == B and precision == "bf16" and tj.get("variant", "i3d") == variant:
                return tj["conv_dram_bytes_per_step"], tj["source"]
    return None, None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import afb200
    from afb200 import parallel, synthetic
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_legs as legs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this repo has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    B = args.batch
    sd = synthetic.synthetic_state_dict(0, args.variant)
    cores = os.cpu_count() or 1

    if args.workload != "clips":                    # one of BASELINE configs 3/4/5 on its own
        eng = afb200.Engine(sd, device=local_rank, max_batch=B, precision=args.precision, variant=args.variant)
        if args.workload == "offline":
            res = legs.offline_leg(eng, B, args.offline_clips, rank, world, dev)
            metric, value, unit = "offline_clips_per_s_32x224x224", res["value"], "clips/s"
        elif args.workload == "live":
            res = legs.live_leg(eng, args.live_streams, args.live_seconds, rank, world, dev)
            metric, value, unit = "live_p50_latency_ms", res["latency_ms_p50"], "ms"
        else:
            res = legs.rgb_leg(sd, args.rgb_batch, rank, world, dev, local_rank)
            metric, value, unit = "rgb_branch_clips_per_s", res["value"], "clips/s"
        if rank == 0:
            _emit(json.dumps({"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
                              "warmup": args.warmup, "higher_is_better": args.workload != "live", "scaling": "strong",
                              "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                              "config": {"workload": res["workload"]}, "result": res}))
        if world > 1:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(cores)
        inputs = make_cpu_clip_inputs(1)
        cpu_reference_step(sd, inputs, 1, args.variant)
        n = 4
        t0 = time.perf_counter()
        cpu_reference_step(sd, inputs, n, args.variant)
        dt = time.perf_counter() - t0
        torch.set_num_threads(1)      # park the OpenMP team: its spinning workers would slow the launch thread below
        cpu_baseline = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "%d clips, batch 1 each (cv2.warpAffine crop + normalise + fp32 forward of the oracle port), after 1 warm-up clip" % n}

    sampler = ClockSampler(local_rank)      # started early so nvidia-smi is already streaming when the load begins
    if rank == 0:
        sampler.start()
    eng = afb200.Engine(sd, device=local_rank, max_batch=B, precision=args.precision, variant=args.variant)
    if args.chunk_front:
        eng.set_option("chunk_front", args.chunk_front)
    if args.chunk_back:
        eng.set_option("chunk_back", args.chunk_back)
    pool, fd, cg, src_bytes, clip_sources = build_gpu_inputs(dev, B, rank)
    n_total = B * world

    def step():
        logits, scores = eng.crop_infer(fd, cg, B)
        if world > 1:
            return logits, parallel.gather_scores(scores, n_total)
        return logits, scores

    # clocks are sampled from the warm-up on (the GPU is under the same load) so that short timed regions
    # still get samples; the timed region itself is bracketed below
    t_load0 = time.time()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    eng.set_option("ridge_x1000", int(1e3 * peaks["bf16_sustained"] * 1e12 / (peaks["hbm_gbs"] * 1e9)))
    eng.set_option("reset_stats", 1)
    eng.set_option("profile_events", 0 if args.no_kernel_events else 1)
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        last_logits, out = step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    ms = parallel.max_over_ranks(ms, dev)
    launches = eng.launch_count - launches0      # close the counting window with the timed region
    eng.set_option("profile_events", 0)
    # nvidia-smi samples every 50 ms: if warm-up + timed region were shorter than ~0.6 s keep the same load running
    # (untimed) until enough samples exist, so `clocks` always describes the GPU under this workload
    extra = 0
    while rank == 0 and time.time() - t_load0 < 0.6 and extra < 200:
        eng.crop_infer(fd, cg, B); torch.cuda.synchronize(); extra += 1      # no collective: rank 0 only
    clocks = sampler.stop(t_load0, time.time()) if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed region + %d untimed steps of the same load" % extra
    tb_ms, tb_flops, tb_n = (eng.get_stat("conv_tensor_bound_" + k) for k in ("ms", "flops", "launches"))
    hb_ms, hb_bytes, hb_n = (eng.get_stat("conv_hbm_bound_" + k) for k in ("ms", "bytes", "launches"))
    conv_ms = eng.get_stat("conv_umma_ms")
    conv_flops = eng.get_stat("conv_umma_flops")
    conv_n = eng.get_stat("conv_umma_launches")
    simt_ms = eng.get_stat("conv_simt_ms")
    conv_alg_bytes = eng.get_stat("conv_bytes")
    k1_ms, k1_n = eng.get_stat("feed_ms"), eng.get_stat("feed_launches")
    value = n_total * args.steps / (ms / 1e3)

    # p50 batch-1 latency (crop + trunk + score on host), rank 0 only; measured right behind the timed region (the
    # host-fed legs below leave the GPU in a lower clock state for a while, which adds ~0.09 ms to a batch-1 pass)
    p50 = p99 = None
    if rank == 0:
        lat = []
        fd1, cg1 = fd[: 32 * 40].contiguous(), cg[:64].contiguous()
        for i in range(55):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lg, sc = eng.crop_infer(fd1, cg1, 1)
            float(sc[0])
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = sorted(lat[5:])
        p50 = lat[len(lat) // 2]
        p99 = lat[min(len(lat) - 1, int(round(0.99 * (len(lat) - 1))))]

    # end to end with HOST inputs, two boundaries, both at --steps:
    #  e2e          crop boundary (matches the reference arm's work: crop/align + pack + classify)
    #  e2e_aligned  ClassifierSvc.infer_scores boundary (pre-aligned u8 clips)
    e2e = e2e_aligned = None
    if not args.no_e2e:
        n_e2e = max(3, args.steps)
        dt, h2d, d2h = legs.e2e_crop_leg(eng, B, n_e2e, 2, rank, dev)
        e2e = {"value": n_total * n_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": n_e2e, "boundary": "crop",
               "api": "afb200.live.FrameRing.put_boxes (af_ring_put_boxes) + af_crop_infer: per step every one of the %d streams "
                      "uploads its 8 new decoded 720p frames (the pixels under the enlarged face box, af_ring_put_boxes) from pinned host memory into the "
                      "device ring, the host computes crop boxes + the similarity fit and the descriptors of the 32-frame "
                      "windows (stride 8), af_crop_infer warps/normalises/classifies, scores are read back; two steps in "
                      "flight" % B}
        dt_a, dt_b = legs.e2e_aligned_leg(eng, B, n_e2e, dev)
        clip_bytes = 32 * 224 * 224 * 3
        e2e_aligned = {"value": n_total * n_e2e / dt_a, "unit": UNIT, "h2d_bytes_per_step": B * clip_bytes,
                       "d2h_bytes_per_step": 8 * B, "steps": n_e2e, "boundary": "aligned clips",
                       "api": "af_submit_u8_host + af_wait (ClassifierSvc.infer_scores_stream: pinned u8 [B,32,224,224,3] -> "
                              "scores on the host, two batches in flight); no crop kernel on this boundary",
                       "blocking_call_value": n_total * n_e2e / dt_b}

    # parity of the very batch that was timed (rank 0): 4 clips back through af_crop_u8 -> CPU oracle
    parity = None
    if rank == 0 and not args.no_parity:
        parity = legs.parity_leg(sd, args.variant, last_logits.cpu().numpy(), clip_sources, [0, B // 3, (2 * B) // 3, B - 1] if B >= 4 else list(range(B)), cores, args.precision)

    # the stock PyTorch (cuDNN) path on the same GPU, N=1 only
    gpu_baseline = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        gpu_baseline = legs.torch_gpu_leg(sd, args.variant, B, dev, "amp_bf16")
        gpu_baseline["tf32"] = legs.torch_gpu_leg(sd, args.variant, B, dev, "tf32")

    # BASELINE configs 3 / 4 / 5, short forms (every rank takes part)
    workloads = None
    if not args.no_extra and args.variant == "i3d" and args.precision == "bf16":
        workloads = {"offline_4096": legs.offline_leg(eng, B, args.offline_clips, rank, world, dev),
                     "live_64_streams": legs.live_leg(eng, args.live_streams, args.live_seconds, rank, world, dev),
                     "rgb_branch_b64": legs.rgb_leg(sd, args.rgb_batch, rank, world, dev, local_rank)}

    launches_t = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(launches_t)
    if rank == 0:
        achieved = conv_flops / (conv_ms * 1e9) if conv_ms > 0 else 0.0
        peak = peaks["bf16_sustained"]
        traffic, traffic_src = _traffic_for(args.variant, B, args.precision)
        floor_ms = FLOPS_PER_CLIP / (peak * 1e12) * 1e3 if args.variant == "i3d" else None
        k1_ms_per_launch = k1_ms / k1_n if k1_n else None
        k1_clips_per_launch = B * args.steps / k1_n if k1_n else None
        roofline_k1 = None
        if k1_ms_per_launch:
            k1_ach = K1_BYTES_PER_CLIP * k1_clips_per_launch / (k1_ms_per_launch * 1e6)       # GB/s
            k1_ach_moved = (src_bytes / B + 32 * 230 * 232 * 8) * k1_clips_per_launch / (k1_ms_per_launch * 1e6)
            roofline_k1 = {"bound": "hbm", "kernel": "crop_kernel (K1: gather + integer-exact warp + normalise + pack)",
                           "achieved": k1_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": k1_ach / peaks["hbm_gbs"],
                           "algorithmic_bytes_per_clip": K1_BYTES_PER_CLIP, "ms_per_launch": k1_ms_per_launch,
                           "clips_per_launch": k1_clips_per_launch,
                           "achieved_bytes_moved": k1_ach_moved,
                           "note": "achieved = 18.7 MB/clip (SURVEY 8d: u8 crops read once + bf16 NCTHW clip written once) x clips per "
                                   "launch / CUDA-event duration of the launch; achieved_bytes_moved counts this run's source boxes and "
                                   "the padded NDHWC4 clip the kernel really writes",
                           "whole_step_form": K1_BYTES_PER_CLIP * value / world / (peaks["hbm_gbs"] * 1e9)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": {"workload": "%s %s, batch %d synthetic 32x224x224 clips per GPU, "
                                       "GPU crop/warp/normalise kernel from 720p frames resident in HBM (BASELINE configs[1])" % (_net_name(args.variant), args.precision, B),
                           "clips_per_step_per_gpu": B, "parallelism": "clip-sharded x%d, scores all-gathered" % world,
                           "l2": "inputs and activations per step (>2 GB) exceed the 126 MB L2; no explicit flush",
                           "weights": "seeded synthetic (no checkpoint ships with the reference)"},
                "e2e": e2e, "e2e_aligned": e2e_aligned, "gpu_launches": int(launches_t.item()),
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": achieved / peak, "traffic": traffic,
                             "traffic_note": "DRAM bytes (read+write) of the %d tcgen05 conv launches of ONE step, from %s" % (int(conv_n / args.steps), traffic_src) if traffic else None,
                             "algorithmic_bytes_per_step": conv_alg_bytes / args.steps,
                             "algorithmic_flops_per_step": conv_flops / args.steps,
                             "kernel": "tcgen05 conv kernels (conv_umma_kernel<64|128|256> + conv_rows_kernel + stem_sweep_kernel + conv_tsweep_kernel), all launches of the timed region",
                             "launches": int(conv_n), "kernel_ms_per_step": conv_ms / args.steps,
                             "share_of_step": conv_ms / ms if ms > 0 else None,
                             "peak_source": "%s bf16_tflops_sustained (MEASURED_PEAKS.json)" % peaks["source"],
                             "simt_conv_ms_per_step": simt_ms / args.steps,
                             "whole_step_frac_of_tensor_roofline": (value / world * FLOPS_PER_CLIP / (peak * 1e12)
                                                                    if args.variant == "i3d" else None),
                             "whole_step_frac_of_burst_roofline": (value / world * FLOPS_PER_CLIP / (peaks["bf16_burst"] * 1e12)
                                                                   if args.variant == "i3d" else None),
                             # the same launches split by the roofline that bounds each (algorithmic FLOP/byte of the
                             # launch vs the ridge peak_tflops / peak_hbm): how close each class runs to ITS limit
                             "classes": {
                                 "tensor_bound": {"launches_per_step": tb_n / args.steps, "ms_per_step": tb_ms / args.steps,
                                                  "achieved": tb_flops / (tb_ms * 1e9) if tb_ms > 0 else None,
                                                  "peak": peak, "unit": "TFLOP/s",
                                                  "frac": tb_flops / (tb_ms * 1e9) / peak if tb_ms > 0 else None},
                                 "hbm_bound": {"launches_per_step": hb_n / args.steps, "ms_per_step": hb_ms / args.steps,
                                               "achieved": hb_bytes / (hb_ms * 1e6) if hb_ms > 0 else None,
                                               "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                               "frac": hb_bytes / (hb_ms * 1e6) / peaks["hbm_gbs"] if hb_ms > 0 else None,
                                               "bytes": "algorithmic (input + output + residual + weights of each launch)"}}},
                "roofline_k1": roofline_k1,
                "cpu_baseline": cpu_baseline, "gpu_baseline": gpu_baseline, "parity": parity, "clocks": clocks,
                "p50_batch1_latency_ms": p50, "p99_batch1_latency_ms": p99,
                "p50_latency_floor_ms": floor_ms, "p50_frac_of_floor": (floor_ms / p50 if (floor_ms and p50) else None),
                "k1_src_bytes_per_clip": src_bytes / B, "workloads": workloads}
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not (parity["max_abs_dlogit"] <= parity["tolerance"] and parity["crop_bit_exact"]):
        sys.stderr.write("bench.py: PARITY FAILED on the timed batch: %s\n" % json.dumps(parity))
        sys.exit(3)


def run_torch_gpu(args):
    """--impl torch_gpu: the stock PyTorch/cuDNN path of the same network on the same B200 (N=1, rank 0)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from afb200 import synthetic
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_legs as legs
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    sd = synthetic.synthetic_state_dict(0, args.variant)
    mode = {"bf16": "amp_bf16", "fp32": "tf32", "tf32": "tf32"}[args.precision]
    res = legs.torch_gpu_leg(sd, args.variant, args.batch, dev, mode, steps=max(1, args.steps), warmup=max(2, args.warmup))
    _emit(json.dumps({"metric": METRIC, "value": res.get("value"), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": res.get("ms_per_step"), "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "impl": "torch_gpu",
                      "config": {"workload": "%s, batch %d, stock PyTorch (%s) on the same GPU, u8 clips resident in HBM" % (_net_name(args.variant), args.batch, mode)},
                      "gpu_baseline": res, "gpu_launches": 0}))


def _net_name(variant):
    return "FTCN-TT (temporal-only ResNet-50 + transformer head)" if variant == "ftcn_tt" else "AltFreezing I3D ResNet-50"


class _QuietStdout:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the
    first communicator when NCCL_DEBUG is set in the environment), so everything but our own line goes to stderr:
    file descriptor 1 points at stderr while the benchmark runs and is restored for the final print."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self._saved, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False


if __name__ == "__main__":
    a = parse()
    with _QuietStdout() as out:
        _emit = out.emit
        if a.impl == "reference":
            run_reference(a)
        elif a.impl == "torch_gpu":
            run_torch_gpu(a)
        else:
            run_ours(a)
