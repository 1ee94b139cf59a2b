"""What perturbs the batch-1 latency: p50 of crop_infer + score read-back measured fresh, after a second stream has
been used, after pinned allocations, and after the host-fed service call."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import afb200, bench
from afb200 import synthetic
dev = torch.device("cuda", 0)
eng = afb200.Engine(synthetic.synthetic_state_dict(0), max_batch=32, precision="bf16")
pool, fd, cg, src_bytes, _ = bench.build_gpu_inputs(dev, 32, 0)
fd1, cg1 = fd[: 32 * 40].contiguous(), cg[:64].contiguous()

def p50(tag):
    lat = []
    for i in range(105):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lg, sc = eng.crop_infer(fd1, cg1, 1)
        float(sc[0])
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = sorted(lat[5:])
    print("%-40s p50 %.3f ms  p10 %.3f  p90 %.3f" % (tag, lat[50], lat[10], lat[90]))

for _ in range(3): eng.crop_infer(fd, cg, 32)
p50("fresh")
s2 = torch.cuda.Stream(device=dev)
with torch.cuda.stream(s2):
    a = torch.zeros(1 << 20, device=dev); a += 1
torch.cuda.synchronize()
p50("after a second stream ran a kernel")
h = torch.empty((32, 32, 224, 224, 3), dtype=torch.uint8).pin_memory()
p50("after a 154 MB pinned allocation")
eng.wait(eng.submit_u8_host_ptr(h.data_ptr(), 32), 32)
p50("after af_submit_u8_host (copy stream + events)")
for _ in range(3): eng.crop_infer(fd, cg, 32)
p50("after 3 more batch-32 steps")
