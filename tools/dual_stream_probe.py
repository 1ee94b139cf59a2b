"""Experiment: two half-batches on disjoint SM halves (two engines, two host threads, two streams) vs one
full-batch engine.  HBM-bound and tensor-bound layers of the two halves can then overlap."""
import os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import afb200
from afb200 import synthetic
B = 32
sd = synthetic.synthetic_state_dict(0)
dev = torch.device("cuda", 0)
u8 = torch.randint(0, 256, (B, 32, 224, 224, 3), dtype=torch.uint8, device=dev)

def bench_single(steps=10):
    eng = afb200.Engine(sd, max_batch=B, precision="bf16")
    for _ in range(3): eng.infer_u8(u8)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps): eng.infer_u8(u8)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    eng.close()
    return B * steps / dt

def bench_dual(limit, nsplit=2, steps=10, offset_ms=0.0):
    engs = [afb200.Engine(sd, max_batch=B // nsplit, precision="bf16") for _ in range(nsplit)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(nsplit)]
    parts = [u8[i * (B // nsplit):(i + 1) * (B // nsplit)].contiguous() for i in range(nsplit)]
    for e in engs: e.set_option("sm_limit", limit)
    def work(i, n):
        with torch.cuda.stream(streams[i]):
            if offset_ms and i: time.sleep(offset_ms * 1e-3 * i)
            for _ in range(n): engs[i].infer_u8(parts[i])
    def run(n):
        th = [threading.Thread(target=work, args=(i, n)) for i in range(nsplit)]
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize()
    run(3)
    t0 = time.perf_counter(); run(steps); dt = time.perf_counter() - t0
    for e in engs: e.close()
    return B * steps / dt

print("single engine, B=32:", round(bench_single(), 1), "clips/s", flush=True)
for limit, nsplit, off in ((74, 2, 0.0), (74, 2, 2.0), (0, 2, 0.0), (96, 2, 2.0), (50, 3, 1.5), (37, 4, 1.0)):
    try:
        print("dual: sm_limit=%d splits=%d offset=%.1fms:" % (limit, nsplit, off), round(bench_dual(limit, nsplit, off=off) if False else bench_dual(limit, nsplit, 10, off), 1), "clips/s", flush=True)
    except Exception as e:
        print("dual failed", limit, nsplit, e, flush=True)
