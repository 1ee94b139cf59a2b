"""Measurement legs used by bench.py (kept out of bench.py so that file stays readable).

Every leg goes through the repo's public host API (afb200.Engine / afb200.live / afb200.crop, i.e. the C ABI of
libafb200.so); the only place a CPU oracle or a library (cuDNN through torch) runs is the explicitly named baseline
and parity legs, never the measured product path.

  e2e_crop_leg      pinned decoded frames on the host -> device frame ring -> af_crop_infer -> scores on the host
                    (the reference hot loop crop_align -> pack -> classify, demo.py:309-328, af_realtime.py:318-360)
  e2e_aligned_leg   pinned aligned u8 clips -> af_submit_u8_host / af_wait -> scores (ClassifierSvc.infer_scores,
                    TEST2.py:151-204)
  parity_leg        clips of the timed batch pulled back through af_crop_u8, CPU oracle forward, logits compared
  torch_gpu_leg     the stock PyTorch path on the same GPU (TEST2.py:136,144,154-156: cudnn.benchmark, autocast,
                    channels_last_3d), i.e. the "existing Blackwell path" the CUDA engine has to beat
  offline_leg       BASELINE config 3: 4096 host-fed clips sharded over the ranks, scores all-gathered
  live_leg          BASELINE config 4: 64 concurrent 30 fps streams, window 32 / stride 8, real-time paced, p50/p99
  rgb_leg           BASELINE config 5: dualrun RGB-branch frame features, bf16, batch 64 over the ranks
"""
import time

import numpy as np
import torch

import afb200
from afb200 import live, parallel, synthetic

H720, W1280 = 720, 1280


def _dist_on():
    return torch.distributed.is_available() and torch.distributed.is_initialized()


def _barrier():
    if _dist_on():
        torch.distributed.barrier()


# --------------------------------------------------------------------------------------------- e2e, crop boundary
class StreamFeeder:
    """`n_streams` face tracks over pinned host frames.  step(i) uploads the 8 new frames of every stream (only the
    pixels its enlarged face box covers) into the device ring and returns the descriptors of the stream's current 32-frame
    window: what a live caller does per stride (af_realtime.py:450-479), batched over the streams of one GPU."""
    SLOTS = 48

    def __init__(self, eng, n_streams, n_steps, seed0=0, stride=8, pinned_frames=40):
        self.eng, self.n, self.stride = eng, n_streams, stride
        t_total = 32 + stride * (n_steps + 1)
        g = torch.Generator().manual_seed(77 + seed0)
        self.host = torch.randint(0, 256, (pinned_frames, H720, W1280, 3), dtype=torch.uint8, generator=g).pin_memory()
        self.ring = live.FrameRing(eng, self.SLOTS * n_streams, H720, W1280)
        tracks = [synthetic.synthetic_track(seed0 + s, t=t_total) for s in range(n_streams)]
        self.det = np.stack([np.stack([b for b, _ in tr]) for tr in tracks])            # [S,t,4] detector boxes
        self.lm5 = np.stack([np.stack([l for _, l in tr]) for tr in tracks])            # [S,t,5,2] frame coordinates
        self.bigs = np.zeros(self.det.shape, np.int64)
        self.copy_stream = torch.cuda.Stream(device=eng.device)
        self.h2d_bytes = 0
        self._frame_bytes = self.host.stride(0)
        self._base = self.host.data_ptr()
        self.put(0, 32)                                                                 # the first window's frames

    def put(self, f0, f1, stream=None):
        """Upload frames [f0,f1) of every stream (the enlarged face box only)."""
        nf = f1 - f0
        new = afb200.get_crop_boxes((H720, W1280), self.det[:, f0:f1].reshape(-1, 4), 0.5).reshape(self.n, nf, 4)
        self.bigs[:, f0:f1] = new
        s_idx = np.repeat(np.arange(self.n), nf)
        f_idx = np.tile(np.arange(f0, f1), self.n)
        slots = s_idx * self.SLOTS + f_idx % self.SLOTS
        src = (s_idx * 5 + f_idx) % self.host.shape[0]
        ptrs = np.uint64(self._base) + src.astype(np.uint64) * np.uint64(self._frame_bytes)
        bx = new.reshape(-1, 4)
        self.ring.put_boxes(slots, ptrs, bx, stream)
        c0, c1 = bx[:, 0] * 3 // 16 * 16, np.minimum((bx[:, 2] * 3 + 15) // 16 * 16, self.host.stride(1))
        self.h2d_bytes += int(((bx[:, 3] - bx[:, 1]) * (c1 - c0)).sum())

    def window(self, i):
        """Descriptors of every stream's window [stride*i, stride*i+32)."""
        f0 = self.stride * i
        bigs = self.bigs[:, f0:f0 + 32]
        lm_rel = self.lm5[:, f0:f0 + 32] - bigs[:, :, None, :2]
        geoms = afb200.clip_geometry_batch(bigs, lm_rel, 224)
        slots = (np.arange(self.n)[:, None] * self.SLOTS + (np.arange(f0, f0 + 32) % self.SLOTS)[None]).reshape(-1)
        buf = self.ring.buf
        fd, cg = afb200.crop.pack_descriptors_ring(buf.data_ptr(), buf.stride(0), buf.stride(1), H720, W1280, slots,
                                                   bigs.reshape(-1, 4), geoms, self.eng.device)
        self.h2d_bytes += fd.numel() + cg.numel()
        return fd, cg


def e2e_crop_leg(eng, B, steps, warmup, rank, dev):
    """clips/s through the crop boundary with HOST inputs: every step uploads the 8 new decoded frames of each of the
    B streams from pinned memory (the pixels under the face box), builds the window descriptors on the host (crop boxes +
    similarity fit, A1/A2), runs af_crop_infer and reads the B scores back; two steps in flight."""
    feeder = StreamFeeder(eng, B, steps + warmup, seed0=1000 * rank)
    main = torch.cuda.current_stream(dev)
    pinned_out = [torch.empty(B, dtype=torch.float32).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    total = np.zeros(1)

    def issue(i):
        # the copy stream may overwrite ring slots last read three windows ago: wait for the step before last
        if i >= 2:
            feeder.copy_stream.wait_event(done[i & 1])
        feeder.put(32 + 8 * (i - 1), 32 + 8 * i, feeder.copy_stream) if i > 0 else None
        copied[i & 1].record(feeder.copy_stream)
        fd, cg = feeder.window(i)
        main.wait_event(copied[i & 1])
        logits, scores = eng.crop_infer(fd, cg, B)
        pinned_out[i & 1].copy_(scores, non_blocking=True)
        done[i & 1].record(main)

    def collect(i):
        done[i & 1].synchronize()
        total[0] += float(pinned_out[i & 1].sum())

    n = 0
    for _ in range(warmup):
        issue(n); collect(n); n += 1
    torch.cuda.synchronize()
    _barrier()
    feeder.h2d_bytes = 0
    t0 = time.perf_counter()
    first = n
    issue(n); n += 1
    for _ in range(steps - 1):
        issue(n)
        collect(n - 1)
        n += 1
    collect(n - 1)
    torch.cuda.synchronize()
    dt = parallel.max_over_ranks(time.perf_counter() - t0, dev)
    assert n - first == steps
    return dt, feeder.h2d_bytes / steps, 4 * B


def e2e_aligned_leg(eng, B, steps, dev):
    """ClassifierSvc boundary: pinned aligned u8 clips -> scores on the host; (pipelined seconds, blocking seconds)."""
    hosts = [torch.empty((B, 32, 224, 224, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    for h_ in hosts:
        h_.random_(0, 256)
    for i in range(2):
        eng.wait(eng.submit_u8_host_ptr(hosts[i].data_ptr(), B), B)
    torch.cuda.synchronize()
    _barrier()
    t0 = time.perf_counter()
    pending = eng.submit_u8_host_ptr(hosts[0].data_ptr(), B)
    for i in range(1, steps):
        nxt = eng.submit_u8_host_ptr(hosts[i & 1].data_ptr(), B)
        eng.wait(pending, B)
        pending = nxt
    eng.wait(pending, B)
    torch.cuda.synchronize()
    dt = parallel.max_over_ranks(time.perf_counter() - t0, dev)
    h_logits = torch.empty(B, dtype=torch.float32).pin_memory()
    h_scores = torch.empty(B, dtype=torch.float32).pin_memory()
    eng.infer_u8_host_ptr(hosts[0].data_ptr(), B, h_logits.data_ptr(), h_scores.data_ptr())
    _barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        eng.infer_u8_host_ptr(hosts[i & 1].data_ptr(), B, h_logits.data_ptr(), h_scores.data_ptr())
    torch.cuda.synchronize()
    dt_block = parallel.max_over_ranks(time.perf_counter() - t0, dev)
    return dt, dt_block


# --------------------------------------------------------------------------------------------- parity of the timed batch
def parity_leg(sd, variant, logits_dev, clip_sources, picked, cores, precision="bf16"):
    """Pull `picked` clips of the timed batch back through af_crop_u8 and run the CPU oracle on them.
    clip_sources: (frames, boxes, geoms) lists of the batch as given to pack_descriptors."""
    from oracle import crop_oracle, ftcn_oracle, i3d_oracle
    net_forward = ftcn_oracle.forward if variant == "ftcn_tt" else i3d_oracle.forward
    frames, boxes, geoms = clip_sources
    torch.set_num_threads(cores)
    got, want = [], []
    crop_exact = True
    for n_, c in enumerate(picked):
        fr = frames[32 * c:32 * c + 32]
        bb = boxes[32 * c:32 * c + 32]
        u8 = afb200.crop.crop_u8(fr, bb, [geoms[c]], 32, 224)[0].cpu().numpy()
        if n_ == 0:          # the crop itself against the integer-exact cv2.warpAffine emulation (one clip: numpy is slow)
            ref_u8 = crop_oracle.crop_align_from_frames([f.cpu().numpy() for f in fr], np.stack(bb), geoms[c][0], geoms[c][1],
                                                        geoms[c][2], 224)
            crop_exact = bool(np.array_equal(u8, ref_u8))
        want.append(float(net_forward(sd, synthetic.normalise_clip(u8))[0, 0]))
        got.append(float(logits_dev[c]))
    torch.set_num_threads(1)
    got, want = np.asarray(got), np.asarray(want)
    med = float(np.median(want))
    decisive = np.abs(want - med) >= 2e-2
    tol = {"bf16": 2e-2, "tf32": 5e-3, "fp32": 1e-3}[precision]      # north_star: bf16 2e-2, fp32 1e-3; tf32: tests/test_tf32_gpu.py
    return {"n": len(picked), "clips": list(picked), "max_abs_dlogit": float(np.abs(got - want).max()), "tolerance": tol,
            "decisions_equal": bool(np.array_equal(got > 0, want > 0)),
            "decisions_equal_recentred": bool(np.array_equal((got > med)[decisive], (want > med)[decisive])),
            "recentred_decisive": int(decisive.sum()), "crop_bit_exact": crop_exact,
            "oracle": "oracle/%s_oracle.py fp32 on the host, on clips read back through af_crop_u8 from the timed batch"
                      % ("ftcn" if variant == "ftcn_tt" else "i3d")}


# --------------------------------------------------------------------------------------------- stock PyTorch on the same GPU
def torch_gpu_leg(sd, variant, B, dev, mode, steps=3, warmup=2):
    """The reference's own GPU configuration of this path (ClassifierSvc.infer_scores, TEST2.py:136,144,151-199):
    u8 clips on the device -> fp32 NCTHW (channels_last_3d) -> normalise -> stock nn-functional forward (cuDNN convs,
    separate BatchNorm / ReLU / pooling kernels) under autocast -> sigmoid.  mode: amp_bf16 | amp_fp16 | tf32 | fp32."""
    from oracle import ftcn_oracle, i3d_oracle
    net_forward = ftcn_oracle.forward if variant == "ftcn_tt" else i3d_oracle.forward
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark = True                                  # DeviceSvc, TEST2.py:136
    tf32 = mode != "fp32"
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    sdd = {k: v.to(dev) for k, v in sd.items()}
    mean = torch.tensor(synthetic.IMAGENET_MEAN, device=dev).view(1, 3, 1, 1, 1) * 255
    std = torch.tensor(synthetic.IMAGENET_STD, device=dev).view(1, 3, 1, 1, 1) * 255
    u8 = torch.randint(0, 256, (B, 32, 224, 224, 3), dtype=torch.uint8, device=dev)
    amp_dtype = {"amp_bf16": torch.bfloat16, "amp_fp16": torch.float16}.get(mode)

    def infer():
        x = u8.to(torch.float32).permute(0, 4, 1, 2, 3).contiguous(memory_format=torch.channels_last_3d)
        x = x.sub(mean).div(std)
        with torch.inference_mode():
            if amp_dtype is not None:
                with torch.autocast("cuda", dtype=amp_dtype):
                    out = net_forward(sdd, x)
            else:
                out = net_forward(sdd, x)
        return torch.sigmoid(out.float())

    try:
        t_w = time.perf_counter()
        for _ in range(warmup):
            infer()
        torch.cuda.synchronize()
        t_w = time.perf_counter() - t_w
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            sc = infer()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res = {"value": B / (ms / 1e3), "unit": "clips/s", "ms_per_step": ms, "batch": B, "mode": mode, "steps": steps,
               "warmup_s": t_w, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
               "what": "stock PyTorch (cuDNN) forward of the same network on the same GPU: u8 clips resident in HBM -> fp32 "
                       "channels_last_3d -> normalise -> conv3d/batch_norm/relu/pool -> sigmoid, cudnn.benchmark on"}
    except Exception as err:                                                # OOM / missing cuDNN engine: report, don't fail the bench
        res = {"value": None, "unit": "clips/s", "mode": mode, "batch": B, "error": "%s: %s" % (type(err).__name__, str(err)[:200])}
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        del sdd, u8
        torch.cuda.empty_cache()
    return res


# --------------------------------------------------------------------------------------------- BASELINE config 3
def offline_leg(eng, B, n_clips, rank, world, dev):
    """batch_eval-style offline scoring (TEST2.py:393-439 flush loop): `n_clips` aligned u8 clips sharded contiguously
    over the ranks, fed from pinned HOST memory in batches of B through the pipelined service call, scores all-gathered."""
    lo, hi = parallel.shard_range(n_clips, rank, world)
    hosts = [torch.empty((B, 32, 224, 224, 3), dtype=torch.uint8).pin_memory() for _ in range(3)]
    for i, h_ in enumerate(hosts):
        h_.random_(0, 256)
    eng.wait(eng.submit_u8_host_ptr(hosts[0].data_ptr(), B), B)
    torch.cuda.synchronize()
    _barrier()
    t0 = time.perf_counter()
    out, pending, k = [], None, 0
    for b0 in range(lo, hi, B):
        nb = min(B, hi - b0)
        t = eng.submit_u8_host_ptr(hosts[k % 3].data_ptr(), nb)
        if pending is not None:
            out.append(eng.wait(pending[0], pending[1])[0])
        pending = (t, nb)
        k += 1
    if pending is not None:
        out.append(eng.wait(pending[0], pending[1])[0])
    local = torch.from_numpy(np.concatenate(out) if out else np.zeros(0, np.float32)).to(dev)
    full = parallel.gather_scores(local, n_clips)
    torch.cuda.synchronize()
    dt = parallel.max_over_ranks(time.perf_counter() - t0, dev)
    return {"workload": "offline scoring, %d aligned clips sharded over %d GPU(s), batch %d, host-fed (pinned u8 clips in, scores "
                        "all-gathered)" % (n_clips, world, B), "clips": n_clips, "value": n_clips / dt, "unit": "clips/s",
            "seconds": dt, "h2d_bytes_per_clip": 32 * 224 * 224 * 3, "scores_gathered": int(full.numel()),
            "reference": "altfreezing/TEST2.py:393-439"}


# --------------------------------------------------------------------------------------------- BASELINE config 4
def live_leg(eng, n_streams, seconds, rank, world, dev, fps=30.0, stride=8):
    """Live-call simulation (af_realtime.py:372-509): `n_streams` concurrent 30 fps streams, sticky to ranks, each new
    frame's face box uploaded from pinned host memory into the rank's device ring as it "arrives" (real-time paced,
    af_ring_put_boxes), a 32-frame
    window scored every `stride` frames per stream, micro-batched per tick.  Latency = window complete -> score on host."""
    mine = [s for s in range(n_streams) if parallel.stream_owner(s, world) == rank]
    n_frames = int(seconds * fps)
    pinned = torch.randint(0, 256, (8, H720, W1280, 3), dtype=torch.uint8).pin_memory()
    pin_base, pin_stride = pinned.data_ptr(), pinned.stride(0)
    SL = 48
    ring = live.FrameRing(eng, SL * max(1, len(mine)), H720, W1280)
    scorers, tracks = {}, {}
    for s in mine:
        scorers[s] = live.LiveScorer(lambda clips: [], 32, stride)
        tr = synthetic.synthetic_track(s, t=n_frames)
        det = np.stack([b for b, _ in tr])
        tracks[s] = (afb200.get_crop_boxes((H720, W1280), det, 0.5), np.stack([l for _, l in tr]))
    if mine:                                              # warm-up: one clip through the fused path
        bigs, lms = tracks[mine[0]]
        warm = [(f, bigs[f], lms[f] - bigs[f][:2][None]) for f in range(32)]
        for f in range(32):
            ring.buf[f].copy_(pinned[f % 8], non_blocking=True)
        fd, cg = live.ring_descriptors(ring, [warm])
        eng.crop_infer(fd, cg, 1)
    torch.cuda.synchronize()
    _barrier()
    lat, late = [], 0
    t0 = time.perf_counter()
    for f in range(n_frames):
        due = t0 + f / fps
        now = time.perf_counter()
        if now < due:
            time.sleep(due - now)
        elif now - due > 1.0 / fps:
            late += 1
        up_slots, up_ptrs, up_boxes = [], [], []
        for li, s in enumerate(mine):
            fs = f - (s % stride)                         # streams join a few frames apart (deterministic phase)
            if fs < 0:
                continue
            bigs, lms = tracks[s]
            slot = li * SL + fs % SL
            up_slots.append(slot)
            up_ptrs.append(pin_base + ((f + s) % 8) * pin_stride)
            up_boxes.append(bigs[fs])
            scorers[s].observe(s, slot, bigs[fs], lms[fs] - bigs[fs][:2][None])
        if up_slots:                                      # this tick's new frames: only their face boxes go to the device
            ring.put_boxes(up_slots, up_ptrs, np.stack(up_boxes))
        pend = [(s, w) for s in mine for (_, w) in scorers[s].pending]
        if pend:
            for s in mine:
                scorers[s].pending.clear()
            for i in range(0, len(pend), eng.max_batch):
                part = pend[i:i + eng.max_batch]
                fd, cg = live.ring_descriptors(ring, [w for _, w in part])
                logits, scores = eng.crop_infer(fd, cg, len(part))
                sc = scores.cpu().numpy()
                done = time.perf_counter()
                for (s, _), v in zip(part, sc):
                    scorers[s].running_scores[s].append(float(v))
                    scorers[s].hyst.update(s, float(v))
                    lat.append((done - due) * 1e3)
    wall = time.perf_counter() - t0
    if _dist_on():
        allv = [None] * world
        torch.distributed.all_gather_object(allv, (lat, late, wall))
        lat = [v for part in allv for v in part[0]]
        late = sum(p[1] for p in allv)
        wall = max(p[2] for p in allv)
    lat = np.asarray(lat)
    return {"workload": "live-call simulation: %d concurrent %.0f fps streams over %d GPU(s), window 32, stride %d, real-time "
                        "paced, the face boxes of the 720p frames uploaded from pinned host memory as they arrive" % (n_streams, fps, world, stride),
            "streams": n_streams, "seconds": seconds, "clips_scored": int(lat.size), "clips_per_s": lat.size / wall,
            "latency_ms_p50": float(np.percentile(lat, 50)) if lat.size else None,
            "latency_ms_p99": float(np.percentile(lat, 99)) if lat.size else None,
            "latency_ms_max": float(lat.max()) if lat.size else None, "late_ticks": int(late),
            "realtime_kept": bool(wall < seconds * 1.05), "latency": "arrival of the window's last frame -> score on the host",
            "reference": "test/af_realtime.py:372-509 (reference p50 5047 ms, BASELINE.md)"}


# --------------------------------------------------------------------------------------------- BASELINE config 5
def rgb_leg(sd, total_batch, rank, world, dev, local_rank, steps=5):
    """dualrun RGB branch (dual_rgb.py:26-44): x [B,T,3,H,W] -> backbone -> [B,16,2048] -> temporal mean, bf16 engine,
    batch `total_batch` split over the ranks, pooled features all-gathered."""
    lo, hi = parallel.shard_range(total_batch, rank, world)
    nb = hi - lo
    clf = afb200.Classifier(precision="bf16", max_batch=min(32, max(1, nb))).to(dev).eval()
    clf.load_state_dict_tolerant(sd)
    backbone = afb200.RGBBackboneB200(clf)
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    x = torch.randn((max(1, nb), 32, 3, 224, 224), device=dev, generator=g).to(torch.bfloat16)

    def step():
        zt = backbone(x)                               # [nb,16,2048]
        pooled = zt.mean(dim=1)
        if _dist_on():
            width = (total_batch + world - 1) // world
            pad = torch.zeros((width, pooled.shape[1]), device=dev)
            pad[:nb] = pooled[:nb]
            out = [torch.empty_like(pad) for _ in range(world)]
            torch.distributed.all_gather(out, pad)
        return pooled

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    _barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1) / steps, dev)
    clf._warped_network.refold()
    return {"workload": "dualrun RGB branch: frames [B,T,3,H,W] bf16 -> per-frame features [B,16,2048] -> temporal mean, batch %d "
                        "over %d GPU(s), pooled features all-gathered" % (total_batch, world),
            "batch_total": total_batch, "value": total_batch / (ms / 1e3), "unit": "clips/s", "ms_per_step": ms, "steps": steps,
            "reference": "dualrun/model/dual_rgb.py:26-44"}
