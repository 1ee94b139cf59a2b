"""Small driver for ncu captures: two trunk passes over B synthetic clips (first is warm-up)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import afb200
from afb200 import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
eng = afb200.Engine(synthetic.synthetic_state_dict(0), max_batch=B, precision="bf16")
u8 = torch.randint(0, 256, (B, 32, 224, 224, 3), dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.infer_u8(u8)
torch.cuda.synchronize()
print("launches", eng.launch_count)
