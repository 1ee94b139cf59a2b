"""Where does batch-1 latency go?  CPU time to enqueue one crop+trunk pass vs time until the score is on the host."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import afb200
from afb200 import synthetic
import bench
dev = torch.device("cuda", 0)
eng = afb200.Engine(synthetic.synthetic_state_dict(0), max_batch=32, precision="bf16")
pool, fd, cg, _, _ = bench.build_gpu_inputs(dev, 2, 0)
fd1, cg1 = fd[: 32 * 40].contiguous(), cg[:64].contiguous()
for _ in range(10):
    eng.crop_infer(fd1, cg1, 1)
torch.cuda.synchronize()
enq, tot = [], []
for _ in range(50):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lg, sc = eng.crop_infer(fd1, cg1, 1)
    t1 = time.perf_counter()
    float(sc[0])
    t2 = time.perf_counter()
    enq.append((t1 - t0) * 1e3); tot.append((t2 - t0) * 1e3)
print("enqueue p50 %.3f ms, total p50 %.3f ms, launches per pass %d" % (sorted(enq)[25], sorted(tot)[25], 0))
# back-to-back passes without host sync: GPU-side time per pass when the CPU runs ahead
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    eng.crop_infer(fd1, cg1, 1)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("20 passes: enqueue %.3f ms/pass, wall %.3f ms/pass" % ((t1 - t0) / 20 * 1e3, (t2 - t0) / 20 * 1e3))
if os.environ.get("AFB200_TIMELINE") == "1":
    eng.crop_infer(fd1, cg1, 1); eng.crop_infer(fd1, cg1, 1)
    torch.cuda.synchronize()
    eng.set_option("dump_timeline", 45)
