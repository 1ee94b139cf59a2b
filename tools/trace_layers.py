"""Per-op timing table of one trunk pass (AFB200_TRACE=1 must be set in the environment)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import afb200
from afb200 import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cf = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
variant = os.environ.get("AFB200_VARIANT", "i3d")
sd = synthetic.synthetic_state_dict(0, variant)
eng = afb200.Engine(sd, max_batch=B, precision="bf16", variant=variant)
if cf: eng.set_option("chunk_front", cf)
if cb: eng.set_option("chunk_back", cb)
u8 = torch.randint(0, 256, (B, 32, 224, 224, 3), dtype=torch.uint8, device="cuda")

eng.infer_u8(u8); torch.cuda.synchronize()
print("---- traced pass B=%d" % B, file=sys.stderr)
eng.infer_u8(u8); torch.cuda.synchronize()
