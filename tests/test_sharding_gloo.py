"""N>1 path on CPU: world_size-2 gloo run of the pair sharding / max-over-ranks / gather logic used by bench.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slam3d_b200 import sharding


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_pairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sources = [np.full((4, 3), i, np.float32) for i in range(n_pairs)]
    targets = [np.full((4, 3), -i, np.float32) for i in range(n_pairs)]

    def fake_align(src, tgt, guesses, params):  # stands in for Context.gicp_align_batch (needs a GPU)
        return [(int(s[0, 0]), int(t[0, 0]), rank) for s, t in zip(src, tgt)]

    lo, hi, res = sharding.align_sharded(fake_align, sources, targets, None, None)
    elapsed = sharding.max_over_ranks(10.0 + rank)
    allres = sharding.gather_to_rank0(res)
    if rank == 0:
        q.put((elapsed, allres))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    for n in (0, 1, 5, 64, 255, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_pairs = 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    elapsed, allres = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert elapsed == 11.0                                   # max over ranks
    assert [r[0] for r in allres] == list(range(n_pairs))    # every pair once, in order
    assert [r[2] for r in allres] == [0, 0, 0, 1, 1, 1, 1]   # contiguous shards
