"""Exchange operators below the C ABI (csrc/exchange.cu): Broadcast and FinalAggregate over the library-owned communicator.
The ranks are emulated by several contexts of this process on the one GPU, each driven by its own thread like a process
would be (qgpu_comm_init_local: a host rendezvous + device copies stand in for NCCL; tests/test_gpu_multigpu.py runs the
same plans over real NCCL when the box has two GPUs).  Every distributed result -- the union of what the ranks return, each
group on exactly one rank -- must equal the single-table plan, which the other tests pin to the oracle."""
import os
import threading

import numpy as np
import pyarrow as pa
import pytest

from qurious_b200 import _lib, tpch
from qurious_b200.distributed import BroadcastJoinAggregate
from qurious_b200.physical.expr import Column, CountAggregateExpr, MaxAggregateExpr, MinAggregateExpr, SumAggregateExpr
from qurious_b200.physical.plan import Broadcast, FinalAggregate, HashAggregate, MemoryTable, Projection, Scan
from tests.cases import check_rows, rows_of

pytestmark = pytest.mark.gpu


def run_ranks(world, work):
    """work(rank, ctx) -> result, one thread per rank over an in-process group; returns the results in rank order."""
    ctxs = [_lib.Context(0) for _ in range(world)]
    _lib.comm_init_local(ctxs)
    out, errs = [None] * world, [None] * world

    def run(r):
        try:
            out[r] = work(r, ctxs[r])
        except Exception as e:      # noqa: BLE001 -- re-raised on the main thread
            errs[r] = e
    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for c in ctxs:
        c.close()
    for e in errs:
        if e is not None:
            raise e
    return out


def shard(mt: MemoryTable, r: int, world: int) -> MemoryTable:
    t = pa.Table.from_batches(mt.data)
    lo, hi = (t.num_rows * r) // world, (t.num_rows * (r + 1)) // world
    return MemoryTable.try_new(mt.schema, t.slice(lo, hi - lo).combine_chunks().to_batches() or
                               [pa.RecordBatch.from_pylist([], schema=mt.schema)])


@pytest.mark.parametrize("world,pruned", [(2, False), (3, True), (8, False), (8, True)])
def test_q3_broadcast_join_final_aggregate(gpu_ctx, world, pruned):
    sf = 0.02
    db = tpch.generate(sf, batch_rows=None)
    single = rows_of(tpch.q3_plan(db).execute(gpu_ctx))

    def work(r, ctx):
        o_sh, l_sh = shard(db.orders, r, world), shard(db.lineitem, r, world)
        cust = MemoryTable.try_new(db.customer.schema, db.customer.data)
        prune = (0, l_sh, l_sh.schema.get_field_index("l_orderkey")) if pruned else None
        bj = BroadcastJoinAggregate(ctx, tpch.q3_build_plan(tpch.Database(sf, cust, o_sh, None)),
                                    lambda b: tpch.q3_probe_plan(b, l_sh), world, prune=prune)
        rounds = [rows_of(bj.execute()) for _ in range(3)]          # learning run, then replays of the learned counts
        strat = bj.last_strategy
        bj.release()
        return rounds, strat
    res = run_ranks(world, work)
    for i in range(3):
        rows = [row for rounds, _ in res for row in rounds[i]]
        keys = [r[0] for r in rows]
        assert len(keys) == len(set(keys)), "a group was returned by two ranks"
        check_rows(f"q3 x{world} round {i}", rows, single, ordered=False)
    assert "final-aggregate[" in res[0][1] and "fused_join_probe_agg" in res[0][1], res[0][1]
    assert ("broadcast[key-range pruned" if pruned else "broadcast[all-gather") in res[0][1], res[0][1]


def test_broadcast_gathers_rows_in_rank_order_with_nulls(gpu_ctx):
    world = 3
    rng = np.random.default_rng(5)
    n = 5000
    a = [None if rng.random() < 0.2 else int(v) for v in rng.integers(-10**12, 10**12, n)]
    d = [None if rng.random() < 0.2 else int(v) for v in rng.integers(0, 20000, n)]
    f = [float(v) for v in rng.standard_normal(n)]
    schema = pa.schema([("a", pa.int64()), ("d", pa.date32()), ("f", pa.float64()), ("z", pa.int32())])
    full = pa.table({"a": pa.array(a, pa.int64()), "d": pa.array(d, pa.int32()).cast(pa.date32()), "f": pa.array(f),
                     "z": pa.array([None] * n, pa.int32())}, schema=schema)
    mt = MemoryTable.try_new(schema, full.to_batches())

    def work(r, ctx):
        sh = shard(mt, r, world) if r != 1 else MemoryTable.try_new(schema, [pa.RecordBatch.from_pylist([], schema=schema)])  # rank 1: empty
        p = Broadcast(Scan(schema, sh, None, None))
        try:
            return [pa.Table.from_batches(p.execute(ctx)) for _ in range(2)]
        finally:
            p.release()
    res = run_ranks(world, work)
    want = pa.concat_tables([pa.Table.from_batches(shard(mt, 0, world).data), pa.Table.from_batches(shard(mt, 2, world).data)]).combine_chunks()
    for per_rank in res:
        for t in per_rank:
            assert t.combine_chunks().equals(want)


def test_final_aggregate_sum_min_max_count(gpu_ctx):
    world = 4
    rng = np.random.default_rng(9)
    n = 40_000
    schema = pa.schema([("k", pa.int64()), ("g", pa.int32()), ("v", pa.int64())])
    full = pa.table({"k": pa.array(rng.integers(0, 700, n)), "g": pa.array(rng.integers(0, 3, n).astype(np.int32)),
                     "v": pa.array(rng.integers(-1000, 1000, n))}, schema=schema)
    mt = MemoryTable.try_new(schema, full.to_batches())
    out = pa.schema([("k", pa.int64()), ("g", pa.int32()), ("s", pa.int64()), ("c", pa.int64()), ("mn", pa.int64()), ("mx", pa.int64())])

    def partial(table):
        K, G, V = Column("k", 0), Column("g", 1), Column("v", 2)
        return HashAggregate(out, Scan(schema, table, None, None), [K, G],
                             [SumAggregateExpr(V, pa.int64()), CountAggregateExpr(V), MinAggregateExpr(V, pa.int64()), MaxAggregateExpr(V, pa.int64())])
    single = rows_of(partial(mt).execute(gpu_ctx))

    def work(r, ctx):
        p = FinalAggregate(partial(shard(mt, r, world)), [0, 1], [(2, "sum"), (3, "sum"), (4, "min"), (5, "max")])
        try:
            return rows_of(p.execute(ctx))
        finally:
            p.release()
    rows = [row for part in run_ranks(world, work) for row in part]
    check_rows("final aggregate", rows, single, ordered=False)


def test_final_aggregate_disjoint_key_ranges_skip_the_exchange(gpu_ctx):
    world = 3
    schema = pa.schema([("k", pa.int64()), ("v", pa.int64())])
    out = pa.schema([("k", pa.int64()), ("s", pa.int64())])

    def work(r, ctx):
        t = MemoryTable.try_new(schema, [pa.record_batch([pa.array(np.arange(1000) % 50 + 100 * r), pa.array(np.arange(1000))], schema=schema)])
        p = FinalAggregate(HashAggregate(out, Scan(schema, t, None, None), [Column("k", 0)], [SumAggregateExpr(Column("v", 1), pa.int64())]),
                           [0], [(1, "sum")])
        try:
            return rows_of(p.execute(ctx)), p.last_strategy()
        finally:
            p.release()
    res = run_ranks(world, work)
    for r, (rows, strat) in enumerate(res):
        assert "disjoint" in strat, strat
        assert sorted(k for k, _ in rows) == [100 * r + i for i in range(50)]


def test_exchange_nodes_are_identity_without_a_communicator(gpu_ctx):
    db = tpch.generate(0.005, batch_rows=None)
    bj = BroadcastJoinAggregate(gpu_ctx, tpch.q3_build_plan(db), lambda b: tpch.q3_probe_plan(b, db.lineitem), 1)
    check_rows("world 1", rows_of(bj.execute()), rows_of(tpch.q3_plan(db).execute(gpu_ctx)), ordered=False)
    bj.release()
