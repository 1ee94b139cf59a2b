"""Driver for ncu captures of K1: the crop/warp/normalise kernel (af_crop_infer from 720p frames, as bench.py's timed
region) and the u8 packer (af_infer_u8), batch B."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import afb200
from afb200 import synthetic
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
eng = afb200.Engine(synthetic.synthetic_state_dict(0), max_batch=B, precision="bf16")
pool, fd, cg, src_bytes, _ = bench.build_gpu_inputs(dev, B, 0)
u8 = torch.randint(0, 256, (B, 32, 224, 224, 3), dtype=torch.uint8, device=dev)
for _ in range(2):
    eng.crop_infer(fd, cg, B)
    eng.infer_u8(u8)
torch.cuda.synchronize()
print("launches", eng.launch_count, "src bytes per clip", src_bytes / B)
