"""BASELINE config 3: batch_eval-style offline scoring — N synthetic aligned clips sharded across the ranks,
batch 32 per step, scores all-gathered.  python tools/offline_eval.py --clips 4096  (torchrun for N>1)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import afb200  # noqa: E402
from afb200 import parallel, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    eng = afb200.Engine(synthetic.synthetic_state_dict(0), device=local, max_batch=a.batch, precision="bf16")
    lo, hi = parallel.shard_range(a.clips, rank, world)
    # 64 distinct synthetic clips per rank, cycled (content does not affect timing); seeded per global index
    base = torch.stack([torch.from_numpy(synthetic.synthetic_clip_u8(lo + i)) for i in range(min(8, hi - lo))]).to(dev)
    eng.infer_u8(base[: min(a.batch, base.shape[0])])
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    out = []
    for b0 in range(lo, hi, a.batch):
        nb = min(a.batch, hi - b0)
        idx = (torch.arange(nb, device=dev) + (b0 - lo)) % base.shape[0]
        logits, scores = eng.infer_u8(base[idx].contiguous())
        out.append(scores)
    local_scores = torch.cat(out) if out else torch.empty(0, device=dev)
    full = parallel.gather_scores(local_scores, a.clips)
    torch.cuda.synchronize()
    dt = parallel.max_over_ranks(time.perf_counter() - t0, dev)
    if rank == 0:
        print(json.dumps({"workload": "offline scoring", "clips": a.clips, "n_gpus": world, "batch": a.batch,
                          "seconds": dt, "clips_per_s": a.clips / dt, "scores_mean": float(full.mean()),
                          "scores_gathered": int(full.numel())}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
