"""Host-side multi-GPU logic on CPU: the ensemble sharding rule, and a world_size-2
gloo run that shards members, does rank-local "work" and reduces statistics the way
bench.py does (max over ranks for time, sum for counts)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

from sayram2d_b200.shard import shard_range, slab_range


@pytest.mark.parametrize("n,world", [(4096, 1), (4096, 2), (4096, 8), (10, 4), (3, 8), (0, 2)])
def test_shard_range_partitions_exactly(n, world):
    ranges = [shard_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a1 >= a0
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n


def test_shard_range_rejects_bad_arguments():
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)
    with pytest.raises(ValueError):
        shard_range(8, 0, 0)
    assert slab_range(16384, 7, 8) == (14336, 16384)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from sayram2d_b200 import fields
    n = 37
    lo, hi = shard_range(n, rank, world)
    a, b = fields.ensemble_scales(np.arange(lo, hi))
    cells = torch.tensor([float((hi - lo) * 6400)], dtype=torch.float64)
    t = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
    asum = torch.tensor([float(a.sum())], dtype=torch.float64)
    dist.all_reduce(cells, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(asum, op=dist.ReduceOp.SUM)
    if rank == 0:
        a_all, _ = fields.ensemble_scales(np.arange(n))
        ok = cells.item() == n * 6400 and abs(t.item() - 0.1 * world) < 1e-12 and abs(asum.item() - a_all.sum()) < 1e-9
        open(out, "w").write("ok" if ok else f"bad {cells.item()} {t.item()} {asum.item()}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "res.txt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
