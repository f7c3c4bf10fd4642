"""N>1 host logic on CPU (gloo, world_size 2): row-range sharding of the generator is consistent with the whole
table, and partial aggregates all-gathered over the process group merge to the whole-table oracle answer."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qurious_b200 import tpch
from qurious_b200.distributed import shard_range

SF = 0.003


def _q6_partial(raw):
    ship, disc, qty, price = (raw.cols[c].numpy() for c in ("l_shipdate", "l_discount", "l_quantity", "l_extendedprice"))
    m = (ship >= tpch.days("1994-01-01")) & (ship < tpch.days("1995-01-01")) & (disc >= 5) & (disc <= 7) & (qty < 2400)
    return int((price[m].astype(object) * disc[m].astype(object)).sum()) if m.any() else 0, int(m.sum())


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = tpch.n_lineitems(SF)
    lo, hi = shard_range(n, rank, world)
    raw = tpch.gen_lineitem(SF, columns=tpch.Q6_COLUMNS, row_range=(lo, hi))
    assert raw.rows == hi - lo
    s, c = _q6_partial(raw)
    # 128-bit sums travel as two 64-bit words (NCCL/gloo have no int128 reduce: all-gather then merge locally)
    mine = torch.tensor([s & (2**63 - 1), s >> 63, c], dtype=torch.int64)
    got = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(got, mine)
    total = sum(int(g[0]) + (int(g[1]) << 63) for g in got)
    rows = sum(int(g[2]) for g in got)
    out[rank] = (total, rows)
    dist.destroy_process_group()


def test_row_range_shards_merge_to_the_whole_table():
    world = 2
    full = tpch.gen_lineitem(SF, columns=tpch.Q6_COLUMNS)
    expect = _q6_partial(full)
    # shards are slices of the same table
    n = tpch.n_lineitems(SF)
    parts = [tpch.gen_lineitem(SF, columns=tpch.Q6_COLUMNS, row_range=shard_range(n, r, world)) for r in range(world)]
    for c in tpch.Q6_COLUMNS:
        assert torch.equal(torch.cat([p.cols[c] for p in parts]), full.cols[c])
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0] == expect and out[1] == expect


@pytest.mark.parametrize("total,world", [(10, 3), (0, 4), (600037902, 8), (7, 8)])
def test_shard_range_partitions(total, world):
    edges = [shard_range(total, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    assert max(h - l for l, h in edges) - min(h - l for l, h in edges) <= 1
