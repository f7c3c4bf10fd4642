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


# ------------------------------------------------------------------------------------------------
# hash repartition / ragged gather protocols over a real process group (gloo, CPU tensors)
# ------------------------------------------------------------------------------------------------
def _repart_worker(rank, world, port, out):
    from qurious_b200.distributed import _dist_all_gather_ragged, _dist_all_to_all, exchange_columns, partition_ids_host
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    n = 1000 + 37 * rank
    k = rng.integers(0, 300, n).astype(np.int64) * 7919
    v = rng.integers(-50, 50, n).astype(np.int32)
    pid = partition_ids_host(k, world)
    order = np.argsort(pid, kind="stable")
    send_rows = [int((pid == p).sum()) for p in range(world)]
    cols = [torch.from_numpy(k[order].copy()).view(torch.uint8), torch.from_numpy(v[order].copy()).view(torch.uint8)]
    outs, recv_rows = exchange_columns(cols, [8, 4], send_rows, _dist_all_to_all)
    rk, rv = outs[0].view(torch.int64).numpy(), outs[1].view(torch.int32).numpy()
    assert len(rk) == len(rv) == sum(recv_rows)
    assert (partition_ids_host(rk, world) == rank).all()
    g, n_all = _dist_all_gather_ragged([torch.from_numpy(rk.copy()).view(torch.uint8)], [8], len(rk), world)
    out[rank] = (sorted(zip(rk.tolist(), rv.tolist())), sorted(zip(k.tolist(), v.tolist())), n_all,
                 sorted(g[0].view(torch.int64).tolist()))
    dist.destroy_process_group()


def _ragged_worker(rank, world, port, out):
    from qurious_b200.distributed import _dist_all_gather_ragged
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = [5, 0, 9][rank]                                     # ragged, one rank empty
    a = (np.arange(n, dtype=np.int64) + 1000 * rank)
    b = (np.arange(n, dtype=np.int32) * 3 + rank)
    c = (np.arange(2 * n, dtype=np.int64) - rank)           # a 16-byte column (two words per row)
    cols = [torch.from_numpy(np.frombuffer(x.tobytes(), dtype=np.uint8).copy()) for x in (a, b, c)]   # byte views, empty included
    g, total = _dist_all_gather_ragged(cols, [8, 4, 16], n, world)
    out[rank] = (total, g[0].view(torch.int64).tolist(), g[1].view(torch.int32).tolist(), g[2].view(torch.int64).tolist())
    dist.destroy_process_group()


def test_ragged_all_gather_packs_columns_in_rank_order_gloo():
    """one collective carries all columns: every rank must see every column's rows in rank order, empty shards included"""
    world = 3
    mgr = mp.Manager()
    out = mgr.dict()
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_ragged_worker, args=(world, port, out), nprocs=world, join=True)
    ns = [5, 0, 9]
    exp_a = sum(([i + 1000 * r for i in range(ns[r])] for r in range(world)), [])
    exp_b = sum(([i * 3 + r for i in range(ns[r])] for r in range(world)), [])
    exp_c = sum(([i - r for i in range(2 * ns[r])] for r in range(world)), [])
    for r in range(world):
        assert out[r] == (14, exp_a, exp_b, exp_c)


def test_hash_repartition_protocol_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_repart_worker, args=(world, port, out), nprocs=world, join=True)
    received = sorted(sum((out[r][0] for r in range(world)), []))
    sent = sorted(sum((out[r][1] for r in range(world)), []))
    assert received == sent                                     # nothing lost, nothing duplicated
    keys = [set(k for k, _ in out[r][0]) for r in range(world)]
    assert not (keys[0] & keys[1])                              # equal keys meet on exactly one rank
    assert out[0][2] == out[1][2] == len(sent) and out[0][3] == out[1][3] == sorted(k for k, _ in sent)
