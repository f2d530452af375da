"""The N>1 path on CPU: world_size-2 gloo processes shard a read set, 'align' their shard, reduce the timing
with MAX and gather the result records on rank 0 - the same helpers bench.py and a multi-GPU driver use."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphaligner_b200 import multi_gpu


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 4, 8):
            b = multi_gpu.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_total, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = multi_gpu.shard_bounds(n_total, world)[rank]
    rec = np.zeros(hi - lo, dtype=multi_gpu.RECORD)
    rec["read"] = np.arange(lo, hi)
    rec["score"] = (np.arange(lo, hi) * 7) % 101
    rec["trace_hash"] = np.arange(lo, hi).astype(np.uint64) * np.uint64(2654435761)
    ms = multi_gpu.reduce_max(dist, 10.0 + rank)
    tot = multi_gpu.reduce_sum(dist, [hi - lo, 1])
    out = multi_gpu.gather_records(dist, rec, n_total, rank, world)
    if rank == 0:
        np.save(os.path.join(tmp, "gathered.npy"), out)
        with open(os.path.join(tmp, "scalars.txt"), "w") as f:
            f.write("%f %f %f" % (ms, tot[0], tot[1]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path):
    n_total, world = 1001, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    out = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    assert list(out["read"]) == list(range(n_total))
    assert list(out["score"]) == [(i * 7) % 101 for i in range(n_total)]
    ms, total, ranks = (float(x) for x in open(os.path.join(str(tmp_path), "scalars.txt")).read().split())
    assert ms == 11.0 and total == n_total and ranks == world
