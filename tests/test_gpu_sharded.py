"""Sharded aggregates (csrc/shard.cu, qurious_b200/distributed.py): N row-range shards emulated on ONE GPU --
the NCCL all-gather is replaced by a concatenation of the state blocks -- must give exactly the rows, values AND
first-occurrence order of the single-table plan."""
import numpy as np
import pyarrow as pa
import pytest
import torch

from oracle import qref
from qurious_b200 import QuriousError, tpch
from qurious_b200.distributed import ShardedAggregate, shard_range
from qurious_b200.physical.expr import (AvgAggregateExpr, Column, CountAggregateExpr, MaxAggregateExpr, MinAggregateExpr,
                                        SumAggregateExpr)
from qurious_b200.physical.plan import HashAggregate, MemoryTable, NoGroupingAggregate, Projection, Scan
from tests.cases import bx, check_rows, lit, rows_of

pytestmark = pytest.mark.gpu


def split_table(t: MemoryTable, world: int):
    full = pa.Table.from_batches(t.data).combine_chunks()
    n = full.num_rows
    shards = []
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        shards.append((lo, MemoryTable.try_new(t.schema, full.slice(lo, hi - lo).to_batches(max_chunksize=1000) or
                                               [pa.record_batch([pa.array([], type=f.type) for f in t.schema], schema=t.schema)])))
    return shards


def run_sharded(ctx, make_plan, table, world, max_groups=64):
    shards = split_table(table, world)
    aggs = [ShardedAggregate(ctx, make_plan(st), lo, world, max_groups=max_groups, all_gather=lambda o, i: None)
            for lo, st in shards]
    states = [a.partial().clone() for a in aggs]
    gathered = torch.cat(states)
    return [a.merge(gathered, world) for a in aggs]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_q1_q6_sharded_equals_single(gpu_ctx, world):
    db = tpch.generate(0.02, batch_rows=None)
    for q in ("q6", "q1"):
        def make(lineitem):
            return getattr(tpch, q + "_plan")(tpch.Database(0.02, None, None, lineitem))
        single = rows_of(make(db.lineitem).execute(gpu_ctx))
        for got in run_sharded(gpu_ctx, make, db.lineitem, world):
            check_rows(f"{q} x{world}", rows_of(got), single, ordered=True)
        check_rows(q + " vs oracle", single, rows_of(qref.execute(make(db.lineitem))), ordered=False)


def test_generic_path_nulls_strings_minmax_sharded(gpu_ctx):
    rng = np.random.default_rng(9)
    n = 6000

    def nullable(vals, t, frac=0.15):
        m = rng.random(n) < frac
        return pa.array([None if d else v for v, d in zip(vals, m)], type=t)
    cols = {"s": nullable([["ab", "BUILDING", "", "sixteen-bytes-key"[:16]][i] for i in rng.integers(0, 4, n)], pa.string()),
            "k": nullable(rng.integers(0, 3, n).tolist(), pa.int32()),
            "v": nullable(rng.integers(-10**9, 10**9, n).tolist(), pa.int64()),
            "f": nullable(rng.normal(0, 1e3, n).tolist(), pa.float64()),
            "d": pa.array([None if x % 11 == 0 else x for x in rng.integers(-10**12, 10**12, n).tolist()], pa.int64())}
    import decimal
    cols["d"] = pa.array([None if x is None else decimal.Decimal(x).scaleb(-2) for x in cols["d"].to_pylist()], pa.decimal128(30, 2))
    schema = pa.schema([(k, v.type) for k, v in cols.items()])
    t = MemoryTable.try_new(schema, [pa.record_batch(list(cols.values()), schema=schema)])
    C = lambda name: Column(name, schema.get_field_index(name))  # noqa: E731
    aggs = [SumAggregateExpr(C("v"), pa.int64()), CountAggregateExpr(C("v")), MinAggregateExpr(C("v"), pa.int64()),
            MaxAggregateExpr(C("f"), pa.float64()), SumAggregateExpr(C("d"), pa.decimal128(30, 2)),
            MinAggregateExpr(C("d"), pa.decimal128(30, 2)), AvgAggregateExpr(C("f"), pa.float64(), pa.float64()),
            AvgAggregateExpr(C("d"), pa.decimal128(30, 2), pa.decimal128(34, 6))]
    out = pa.schema([("s", pa.string()), ("k", pa.int32())] + [(f"a{i}", a.return_type) for i, a in enumerate(aggs)])

    def make(tab):
        return HashAggregate(out, Scan(schema, tab, None, bx(C("v"), "Gt", lit(-9 * 10**8))), [C("s"), C("k")], aggs)
    single = rows_of(make(t).execute(gpu_ctx))
    ref = rows_of(qref.execute(make(t)))
    assert len(single) == len(ref)
    for got in run_sharded(gpu_ctx, make, t, 4):
        rows = rows_of(got)
        assert len(rows) == len(single)
        for a, b in zip(rows, single):          # same order, integer/decimal columns bit-exact
            assert a[:8] == b[:8] and a[9] == b[9], (a, b)
            assert (a[8] is None and b[8] is None) or abs(a[8] - b[8]) <= 1e-12 * abs(b[8])

    def make_ng(tab):
        return NoGroupingAggregate(pa.schema([(f"a{i}", a.return_type) for i, a in enumerate(aggs)]), Scan(schema, tab, None, None), aggs)
    single = rows_of(make_ng(t).execute(gpu_ctx))
    for got in run_sharded(gpu_ctx, make_ng, t, 3):
        a, b = rows_of(got)[0], single[0]
        assert a[:6] == b[:6] and a[7] == b[7]
        assert abs(a[6] - b[6]) <= 1e-12 * abs(b[6])


def test_too_many_groups_is_an_error(gpu_ctx):
    n = 5000
    schema = pa.schema([("k", pa.int64()), ("v", pa.int64())])
    t = MemoryTable.try_new(schema, [pa.record_batch([pa.array(np.arange(n)), pa.array(np.ones(n, dtype=np.int64))], schema=schema)])

    def make(tab):
        return HashAggregate(pa.schema([("k", pa.int64()), ("c", pa.int64())]), Scan(schema, tab, None, None), [Column("k", 0)],
                             [CountAggregateExpr(Column("v", 1))])
    with pytest.raises(QuriousError):
        run_sharded(gpu_ctx, make, t, 2, max_groups=64)


# ------------------------------------------------------------------------------------------------
# The fused protocol (qgpu_plan_execute_sharded*): scan kernel + ONE epilogue kernel that stores the state block into
# every peer's symmetric buffer, waits on the peers' flags, merges and finalises.  Ranks are emulated by several
# contexts of this process on the one GPU (qgpu_comm_init_local: same kernels, peer pointers instead of CUDA IPC),
# each driven by its own thread like a process would be.
# ------------------------------------------------------------------------------------------------
def run_fused_sharded(make_plan, table, world, max_groups=64, rounds=2):
    import os
    import threading
    from qurious_b200 import _lib
    os.environ.setdefault("QGPU_PEER_TIMEOUT_MS", "8000")
    ctxs = [_lib.Context(0) for _ in range(world)]
    _lib.comm_init_local(ctxs)
    shards = split_table(table, world)
    out, errs, keep = [[] for _ in range(world)], [None] * world, [None] * world

    def work(r):
        try:
            lo, st = shards[r]
            plan = make_plan(st)
            agg = ShardedAggregate(ctxs[r], plan, lo, world, max_groups=max_groups)
            assert agg.fused
            keep[r] = (plan, st)
            for i in range(rounds):
                if i % 2 == 0:
                    out[r].append(rows_of(agg.execute()))
                else:                                   # asynchronous device-resident variant
                    t = agg.execute_device(wait=False)
                    t.wait()
                    out[r].append(rows_of([t.to_batch()] if t.num_rows else []))
                    t.free()
            out[r].append(plan.last_strategy())
        except Exception as e:      # noqa: BLE001 -- re-raised on the main thread
            errs[r] = e
    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for r in range(world):
        if keep[r] is not None:
            keep[r][0].release()
            if keep[r][1]._dev is not None:
                keep[r][1]._dev.free()
    for c in ctxs:
        c.close()
    for e in errs:
        if e is not None:
            raise e
    return out


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_peer_exchange_q1_q6(gpu_ctx, world):
    db = tpch.generate(0.02, batch_rows=None)
    for q in ("q6", "q1"):
        def make(lineitem):
            return getattr(tpch, q + "_plan")(tpch.Database(0.02, None, None, lineitem))
        single = rows_of(make(db.lineitem).execute(gpu_ctx))
        for per_rank in run_fused_sharded(make, db.lineitem, world):
            assert "single-CTA epilogue[peer exchange + merge over %d ranks]" % world in per_rank[-1], per_rank[-1]
            for got in per_rank[:-1]:
                check_rows(f"{q} fused x{world}", got, single, ordered=True)


def test_fused_peer_exchange_generic_sources(gpu_ctx):
    """Plans the dense fused kernel refuses (NULLs, Utf8 + Int32 keys, Float64 MAX, wide decimals) pack their generic
    accumulators into the same state block; the exchange / merge / finalise kernel is the same."""
    rng = np.random.default_rng(19)
    n = 5000

    def nullable(vals, t, frac=0.15):
        m = rng.random(n) < frac
        return pa.array([None if d else v for v, d in zip(vals, m)], type=t)
    import decimal
    cols = {"s": nullable([["ab", "BUILDING", "", "sixteen-bytes-key"[:16]][i] for i in rng.integers(0, 4, n)], pa.string()),
            "k": nullable(rng.integers(0, 3, n).tolist(), pa.int32()),
            "v": nullable(rng.integers(-10**9, 10**9, n).tolist(), pa.int64()),
            "f": nullable(rng.normal(0, 1e3, n).tolist(), pa.float64()),
            "d": pa.array([None if x % 11 == 0 else decimal.Decimal(x).scaleb(-2) for x in rng.integers(-10**12, 10**12, n).tolist()],
                          pa.decimal128(30, 2))}
    schema = pa.schema([(k, v.type) for k, v in cols.items()])
    t = MemoryTable.try_new(schema, [pa.record_batch(list(cols.values()), schema=schema)])
    C = lambda name: Column(name, schema.get_field_index(name))  # noqa: E731
    aggs = [SumAggregateExpr(C("v"), pa.int64()), CountAggregateExpr(C("v")), MinAggregateExpr(C("v"), pa.int64()),
            MaxAggregateExpr(C("f"), pa.float64()), SumAggregateExpr(C("d"), pa.decimal128(30, 2)),
            MinAggregateExpr(C("d"), pa.decimal128(30, 2)), AvgAggregateExpr(C("f"), pa.float64(), pa.float64()),
            AvgAggregateExpr(C("d"), pa.decimal128(30, 2), pa.decimal128(34, 6))]
    out = pa.schema([("s", pa.string()), ("k", pa.int32())] + [(f"a{i}", a.return_type) for i, a in enumerate(aggs)])

    def make(tab):
        return HashAggregate(out, Scan(schema, tab, None, bx(C("v"), "Gt", lit(-9 * 10**8))), [C("s"), C("k")], aggs)
    single = rows_of(make(t).execute(gpu_ctx))
    for per_rank in run_fused_sharded(make, t, 3):
        for rows in per_rank[:-1]:
            assert len(rows) == len(single)
            for a, b in zip(rows, single):          # same order, integer/decimal columns bit-exact
                assert a[:8] == b[:8] and a[9] == b[9], (a, b)
                assert (a[8] is None and b[8] is None) or abs(a[8] - b[8]) <= 1e-12 * abs(b[8])

    def make_ng(tab):
        return NoGroupingAggregate(pa.schema([(f"a{i}", a.return_type) for i, a in enumerate(aggs)]), Scan(schema, tab, None, None), aggs)
    single = rows_of(make_ng(t).execute(gpu_ctx))
    for per_rank in run_fused_sharded(make_ng, t, 2):
        for rows in per_rank[:-1]:
            a, b = rows[0], single[0]
            assert a[:6] == b[:6] and a[7] == b[7]
            assert abs(a[6] - b[6]) <= 1e-12 * abs(b[6])


def test_fused_peer_exchange_empty_shard_and_typed_sentinels(gpu_ctx):
    """Ungrouped MIN / MAX where no row qualifies on ANY shard keep the typed start values; a predicate that only one
    shard satisfies merges with the other shards' empty states."""
    n = 4000
    rng = np.random.default_rng(5)
    schema = pa.schema([("i", pa.int32()), ("p", pa.decimal128(15, 2)), ("v", pa.int64())])
    import decimal
    t = MemoryTable.try_new(schema, [pa.record_batch(
        [pa.array(rng.integers(-100, 100, n).astype(np.int32)),
         pa.array([decimal.Decimal(int(x)).scaleb(-2) for x in rng.integers(0, 10**7, n)], pa.decimal128(15, 2)),
         pa.array(np.arange(n, dtype=np.int64))], schema=schema)])
    osch = pa.schema([("mn_i", pa.int32()), ("mx_p", pa.decimal128(15, 2)), ("s", pa.int64()), ("c", pa.int64())])
    for pred in (bx(Column("v", 2), "Lt", lit(-1)), bx(Column("v", 2), "Lt", lit(700))):
        def make(tab):
            return NoGroupingAggregate(osch, Scan(schema, tab, None, pred),
                                       [MinAggregateExpr(Column("i", 0), pa.int32()), MaxAggregateExpr(Column("p", 1), pa.decimal128(15, 2)),
                                        SumAggregateExpr(Column("v", 2), pa.int64()), CountAggregateExpr(Column("v", 2))])
        single = rows_of(make(t).execute(gpu_ctx))
        check_rows("oracle", single, rows_of(qref.execute(make(t))), ordered=True)
        for per_rank in run_fused_sharded(make, t, 4):
            for rows in per_rank[:-1]:
                check_rows("fused empty shards", rows, single, ordered=True)


def test_fused_too_many_groups_is_an_error_on_every_rank(gpu_ctx):
    n = 3000
    schema = pa.schema([("k", pa.int64()), ("v", pa.int64())])
    t = MemoryTable.try_new(schema, [pa.record_batch([pa.array(np.arange(n) % 1500), pa.array(np.ones(n, dtype=np.int64))], schema=schema)])

    def make(tab):
        return HashAggregate(pa.schema([("k", pa.int64()), ("c", pa.int64())]), Scan(schema, tab, None, None), [Column("k", 0)],
                             [CountAggregateExpr(Column("v", 1))])
    with pytest.raises(QuriousError):
        run_fused_sharded(make, t, 2, max_groups=64, rounds=1)
