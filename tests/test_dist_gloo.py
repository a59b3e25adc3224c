"""CPU, world_size 2 over gloo: the prompt-sharding / final-gather plumbing of the multi-GPU path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmada_b200.dist import gather_rows, prompt_seed, shard_range, sharded_generate


def test_shard_range_partitions():
    for n in (1, 7, 8, 9, 64):
        for w in (1, 2, 3, 8):
            cover = []
            for r in range(w):
                lo, hi = shard_range(n, r, w)
                cover += list(range(lo, hi))
            assert cover == list(range(n))
    assert len({prompt_seed(1, i) for i in range(1000)}) == 1000


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def fn(lo, hi):          # stands in for t2i_generate on this rank's prompts: row i depends on i only
        return torch.stack([torch.arange(4, dtype=torch.int64) + 100 * i for i in range(lo, hi)]) if hi > lo \
            else torch.zeros((0, 4), dtype=torch.int64)
    out = sharded_generate(fn, n)
    ok = torch.equal(out, torch.stack([torch.arange(4, dtype=torch.int64) + 100 * i for i in range(n)]))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_generate_world2():
    for n in (8, 5):
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        assert sorted(res) == [(0, True), (1, True)]


def test_gather_rows_single_process_is_identity():
    t = torch.arange(6).view(3, 2)
    assert gather_rows(t, 3) is t
