"""Throughput of the FTCN-TT plugin path on one B200: u8 aligned clips resident in HBM ->
normalise/pack -> temporal-only trunk -> transformer head -> scores.  Prints one JSON line.  (The CPU baseline of
this plugin is timed by `bench.py --variant ftcn_tt`, the one place outside tests/ that may run oracle/.)"""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import afb200
from afb200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
sd = synthetic.synthetic_state_dict(0, "ftcn_tt")
eng = afb200.Engine(sd, max_batch=a.batch, precision="bf16", variant="ftcn_tt")
u8 = torch.randint(0, 256, (a.batch, 32, 224, 224, 3), dtype=torch.uint8, device="cuda")
for _ in range(a.warmup):
    eng.infer_u8(u8)
torch.cuda.synchronize()
n0 = eng.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    lg, sc = eng.infer_u8(u8)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
lat = []
for _ in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lg1, sc1 = eng.infer_u8(u8[:1]); float(sc1[0])
    lat.append((time.perf_counter() - t0) * 1e3)
out = {"metric": "clips_per_s_32x224x224", "variant": "ftcn_tt", "value": a.batch / ms * 1e3, "unit": "clips/s",
       "ms_per_step": ms, "batch": a.batch, "steps": a.steps, "dtype": "bf16", "data": "synthetic",
       "gpu_launches": eng.launch_count - n0 - 0, "p50_batch1_latency_ms": sorted(lat[5:])[len(lat[5:]) // 2]}
print(json.dumps(out))
