"""world_size-2 gloo test of the sharding + score gather used for N>1 GPUs."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    sys.path.insert(0, ROOT)
    from afb200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_range(n_items, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32) * 0.5 + 1.0
    full = parallel.gather_scores(local, n_items)
    mx = parallel.max_over_ranks(float(rank + 1), torch.device("cpu"))
    q.put((rank, full.tolist(), mx))
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    for n_items in (7, 8):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        want = [i * 0.5 + 1.0 for i in range(n_items)]
        for rank, full, mx in res:
            assert full == want
            assert mx == 2.0


def test_shard_range_covers_everything():
    from afb200 import parallel
    for n in (0, 1, 5, 4096, 4097):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.stream_owner(13, 8) == 5
