"""Hash repartition and gather-merge on ONE GPU: the N ranks are emulated, the NCCL collectives are
replaced by slicing / concatenating the same device buffers.  The reference is single-process, so the checks are
(a) partition ids bit-equal to the host mirror of the partition function, (b) the rows are a permutation of the
input, (c) the distributed result equals the single-table plan (which the other GPU tests pin to the oracle)."""
import decimal

import numpy as np
import pyarrow as pa
import pytest
import torch

from oracle import qref
from qurious_b200 import QuriousError, tpch
from qurious_b200.distributed import (BroadcastJoinAggregate, GatherMergeAggregate, column_bytes_tensor, merge_spec_of, partition_ids_host,
                                      shard_range, table_from_tensors)
from qurious_b200.physical.expr import (AvgAggregateExpr, Column, CountAggregateExpr, MaxAggregateExpr, MinAggregateExpr,
                                        SumAggregateExpr)
from qurious_b200.physical.plan import HashAggregate, MemoryTable, Scan
from tests.cases import check_rows, rows_of

pytestmark = pytest.mark.gpu

SCHEMA = pa.schema([("k", pa.int64()), ("v", pa.int64()), ("f", pa.float64()), ("d", pa.date32()),
                    ("w", pa.decimal128(30, 2))])


def make_table(n, groups, seed=3):
    rng = np.random.default_rng(seed)
    with np.errstate(over="ignore"):
        k = (rng.integers(0, max(groups, 1), n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)).view(np.int64)
    cols = [pa.array(k), pa.array(rng.integers(-10**6, 10**6, n).astype(np.int64)), pa.array(rng.random(n)),
            pa.array(rng.integers(0, 20000, n).astype(np.int32), pa.date32()),
            pa.array([decimal.Decimal(int(x)).scaleb(-2) for x in rng.integers(-10**15, 10**15, n)], pa.decimal128(30, 2))]
    return MemoryTable.try_new(SCHEMA, [pa.record_batch(cols, schema=SCHEMA)])


def groupby_plan(table):
    K, V, F, W = (Column(n, SCHEMA.get_field_index(n)) for n in ("k", "v", "f", "w"))
    aggs = [SumAggregateExpr(V, pa.int64()), CountAggregateExpr(V), MinAggregateExpr(V, pa.int64()), MaxAggregateExpr(V, pa.int64()),
            AvgAggregateExpr(F, pa.float64(), pa.float64()), SumAggregateExpr(W, pa.decimal128(38, 2))]
    out = pa.schema([("k", pa.int64())] + [(f"a{i}", a.return_type) for i, a in enumerate(aggs)])
    return HashAggregate(out, Scan(SCHEMA, table, None, None), [K], aggs)


@pytest.mark.parametrize("n,n_parts", [(0, 2), (1, 1), (5000, 2), (100_000, 3), (100_000, 8), (33_333, 64)])
def test_hash_partition_rows_and_ids(gpu_ctx, n, n_parts):
    t = make_table(n, max(n // 10, 1))
    dev = t.device_table(gpu_ctx)
    parted, offs = dev.hash_partition(0, n_parts)
    assert offs[0] == 0 and offs[-1] == n and all(a <= b for a, b in zip(offs, offs[1:]))
    got = parted.to_batch() if n else None
    parted.free()
    if n == 0:
        return
    keys = got.column(0).to_numpy()
    pid = partition_ids_host(keys, n_parts)
    for p in range(n_parts):
        assert (pid[offs[p]:offs[p + 1]] == p).all()
    src = t.data[0]
    a = sorted(zip(*[src.column(i).to_pylist() for i in range(len(SCHEMA))]))
    b = sorted(zip(*[got.column(i).to_pylist() for i in range(len(SCHEMA))]))
    assert a == b            # the same rows, bit for bit (floats included: they are only moved)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_repartitioned_groupby_equals_single(gpu_ctx, world):
    n = 60_000
    t = make_table(n, 7000, seed=world)
    full = pa.Table.from_batches(t.data).combine_chunks()
    single = rows_of(groupby_plan(t).execute(gpu_ctx))
    check_rows("single vs oracle", single, rows_of(qref.execute(groupby_plan(t))), ordered=False)
    # every emulated rank partitions its row range; "all-to-all" = slice p of every rank's buffers goes to rank p
    parted = []
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        st = MemoryTable.try_new(SCHEMA, full.slice(lo, hi - lo).to_batches())
        pt, offs = st.device_table(gpu_ctx).hash_partition(0, world)
        parted.append((pt, offs, [column_bytes_tensor(pt, c) for c in range(len(SCHEMA))]))
    rows, seen = [], set()
    for p in range(world):
        cols, n_recv = [], 0
        for c in range(len(SCHEMA)):
            w = parted[0][2][c][1]
            cols.append(torch.cat([pt_cols[c][0][offs[p] * w:offs[p + 1] * w] for _, offs, pt_cols in parted]))
        n_recv = sum(offs[p + 1] - offs[p] for _, offs, _ in parted)
        local = table_from_tensors(gpu_ctx, SCHEMA, cols, n_recv)
        got = rows_of(groupby_plan(MemoryTable.from_device_table(local)).execute(gpu_ctx)) if n_recv else []
        keys = {r_[0] for r_ in got}
        assert not (keys & seen)          # groups are rank-disjoint after the exchange
        seen |= keys
        rows += got
    check_rows(f"repartitioned x{world}", rows, single, ordered=False)


def test_narrowed_decimal_has_no_arrow_buffer(gpu_ctx):
    schema = pa.schema([("k", pa.int64()), ("p", pa.decimal128(15, 2))])
    t = MemoryTable.try_new(schema, [pa.record_batch([pa.array([1, 2, 3]), pa.array([decimal.Decimal("1.25")] * 3, pa.decimal128(15, 2))],
                                                     schema=schema)])
    dev = t.device_table(gpu_ctx)
    with pytest.raises(QuriousError):
        dev.column_device_buffer(1)


@pytest.mark.parametrize("world", [2, 4])
def test_q3_gather_merge_equals_single(gpu_ctx, world):
    sf = 0.02
    db = tpch.generate(sf, batch_rows=None)
    single = rows_of(tpch.q3_plan(db).execute(gpu_ctx))
    check_rows("q3 vs oracle", single, rows_of(qref.execute(tpch.q3_plan(db))), ordered=False)
    full = pa.Table.from_batches(db.lineitem.data).combine_chunks()
    n = full.num_rows
    plans = []
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        li = MemoryTable.try_new(db.lineitem.schema, full.slice(lo, hi - lo).to_batches())
        plans.append(tpch.q3_plan(tpch.Database(sf, db.customer, db.orders, li)))
    assert merge_spec_of(plans[0]) == ([0, 2, 3], [(1, "sum")])
    locals_ = []
    for p in plans:                               # phase 1: every rank's local result, kept as byte tensors
        d = p.execute_device(gpu_ctx)
        locals_.append(([column_bytes_tensor(d, c)[0].clone() for c in range(len(p.schema))], d.num_rows))
        d.free()
    assert sum(nr for _, nr in locals_) >= len(single)

    def fake_gather(cols, widths, n_rows, w):
        return [torch.cat([l[0][c] for l in locals_]) for c in range(len(cols))], sum(l[1] for l in locals_)
    for p in plans:                               # phase 2: every rank merges the gathered rows identically
        got = rows_of(GatherMergeAggregate(gpu_ctx, p, world, all_gather_ragged=fake_gather).execute())
        check_rows(f"q3 gather-merge x{world}", got, single, ordered=False)


# ------------------------------------------------------------------------------------------------
# exchange (P2P scatter) with the ranks emulated inside one process on one GPU
# ------------------------------------------------------------------------------------------------
def groupby_plan_i64(table):
    """the config-4 shape: SUM/COUNT/MIN/MAX(v), AVG(f) by k -- 64-bit accumulators only (radix / exchange eligible)"""
    K, V, F = (Column(n, SCHEMA.get_field_index(n)) for n in ("k", "v", "f"))
    aggs = [SumAggregateExpr(V, pa.int64()), CountAggregateExpr(V), MinAggregateExpr(V, pa.int64()), MaxAggregateExpr(V, pa.int64()),
            AvgAggregateExpr(F, pa.float64(), pa.float64())]
    out = pa.schema([("k", pa.int64())] + [(f"a{i}", a.return_type) for i, a in enumerate(aggs)])
    return HashAggregate(out, Scan(SCHEMA, table, None, None), [K], aggs)


class _Emu:
    """Collectives of `world` emulated ranks that run their stages in lock-step on one GPU."""

    def __init__(self, world):
        self.world, self.box = world, {}

    def put(self, name, rank, t):
        self.box.setdefault(name, {})[rank] = t.clone()

    def cat(self, name):
        return torch.cat([self.box[name][r] for r in range(self.world)])


def run_exchange_emulated(gpu_ctx, table, world, make_plan, key_cols=1):
    from qurious_b200.distributed import ExchangeGroupBy
    full = pa.Table.from_batches(table.data).combine_chunks()
    n = full.num_rows
    ex = []
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        st = MemoryTable.try_new(table.schema, full.slice(lo, hi - lo).to_batches())
        ex.append(ExchangeGroupBy(gpu_ctx, make_plan(st), world, r, collectives=object()))
    emu = _Emu(world)
    stats = [e.keystats() for e in ex]
    gstats = [(min if i % 2 == 0 else max)(s_[i] for s_ in stats) for i in range(len(stats[0]))]
    for r, e in enumerate(ex):
        sk, ok = e.sketch(gstats)
        assert ok
        emu.put("sketch", r, sk)
    torch.cuda.synchronize()
    gathered = emu.cat("sketch").cpu()
    oks = []
    for r, e in enumerate(ex):
        h, ok = e.prepare(gathered)
        oks.append(ok)
        emu.put("handles", r, h)
    assert len(set(oks)) == 1          # the eligibility decision is collective
    if not oks[0]:
        return None
    all_h = emu.cat("handles")
    for e in ex:
        e.scatter(all_h)
    rows, seen = [], set()
    for e in ex:
        assert e.finish() == 0
        got = rows_of(e.plan.execute(gpu_ctx))
        assert "exchange[rank" in e.plan.last_strategy(), e.plan.last_strategy()
        keys = {tuple(r_[:key_cols]) for r_ in got}
        assert not (keys & seen)
        seen |= keys
        rows += got
    return rows


@pytest.mark.parametrize("world,n,groups", [(2, 60_000, 7000), (3, 50_000, 50_000), (8, 200_000, 30_000), (4, 9, 4), (1, 20_000, 500)])
def test_exchange_groupby_equals_single(gpu_ctx, world, n, groups):
    t = make_table(n, groups, seed=world + 10)
    single = rows_of(groupby_plan_i64(t).execute(gpu_ctx))
    check_rows("single vs oracle", single, rows_of(qref.execute(groupby_plan_i64(t))), ordered=False)
    rows = run_exchange_emulated(gpu_ctx, t, world, groupby_plan_i64)
    assert rows is not None
    check_rows(f"exchange x{world}", rows, single, ordered=False)
    # a second execution re-uses the receive buffers
    rows2 = run_exchange_emulated(gpu_ctx, t, world, groupby_plan_i64)
    check_rows(f"exchange x{world} again", rows2, single, ordered=False)


@pytest.mark.parametrize("world", [2, 3])
def test_q3_broadcast_build_equals_single(gpu_ctx, world):
    """orders AND lineitem row-range sharded; J1 per shard, its rows all-gathered, J2 + aggregate per lineitem shard."""
    sf = 0.02
    db = tpch.generate(sf, batch_rows=None)
    single = rows_of(tpch.q3_plan(db).execute(gpu_ctx))
    orders = pa.Table.from_batches(db.orders.data).combine_chunks()
    items = pa.Table.from_batches(db.lineitem.data).combine_chunks()
    shards = []
    for r in range(world):
        olo, ohi = shard_range(orders.num_rows, r, world)
        llo, lhi = shard_range(items.num_rows, r, world)
        shards.append((MemoryTable.try_new(db.orders.schema, orders.slice(olo, ohi - olo).to_batches()),
                       MemoryTable.try_new(db.lineitem.schema, items.slice(llo, lhi - llo).to_batches())))
    # phase 1: every rank's J1 rows; phase 2: probe results; the emulated "all-gathers" concatenate them in rank order
    builds = []
    for o_sh, _ in shards:
        d = tpch.q3_build_plan(tpch.Database(sf, db.customer, o_sh, None)).execute_device(gpu_ctx)
        builds.append(([column_bytes_tensor(d, c)[0].clone() for c in range(3)], d.num_rows))
        d.free()
    gathered_build = table_from_tensors(gpu_ctx, tpch.Q3_BUILD_SCHEMA, [torch.cat([b[0][c] for b in builds]) for c in range(3)],
                                        sum(b[1] for b in builds))
    probes = []
    for _, l_sh in shards:
        p = tpch.q3_probe_plan(MemoryTable.from_device_table(gathered_build), l_sh)
        d = p.execute_device(gpu_ctx)
        probes.append(([column_bytes_tensor(d, c)[0].clone() for c in range(len(p.schema))], d.num_rows))
        d.free()

    def fake_gather(cols, widths, n_rows, w):
        src = builds if len(cols) == 3 else probes
        return [torch.cat([b[0][c] for b in src]) for c in range(len(cols))], sum(b[1] for b in src)
    for o_sh, l_sh in shards:
        bj = BroadcastJoinAggregate(gpu_ctx, tpch.q3_build_plan(tpch.Database(sf, db.customer, o_sh, None)),
                                    lambda b, l_sh=l_sh: tpch.q3_probe_plan(b, l_sh), world, all_gather_ragged=fake_gather)
        check_rows(f"q3 broadcast x{world}", rows_of(bj.execute()), single, ordered=False)
        assert "broadcast-build" in bj.last_strategy


def test_exchange_two_keys_and_predicate(gpu_ctx):
    """keys (d, k) packed over the GLOBAL value ranges (the shards see different minima), pushed-down predicate"""
    from tests.cases import bx, lit
    t = make_table(40_000, 900, seed=21)
    K, V, F, D = (Column(n, SCHEMA.get_field_index(n)) for n in ("k", "v", "f", "d"))

    def plan(table):
        aggs = [SumAggregateExpr(V, pa.int64()), CountAggregateExpr(V), AvgAggregateExpr(F, pa.float64(), pa.float64())]
        out = pa.schema([("d", pa.date32()), ("v", pa.int64())] + [(f"a{i}", a.return_type) for i, a in enumerate(aggs)])
        return HashAggregate(out, Scan(SCHEMA, table, None, bx(V, "Gt", lit(-900_000))), [D, V], aggs)
    single = rows_of(plan(t).execute(gpu_ctx))
    check_rows("single vs oracle", single, rows_of(qref.execute(plan(t))), ordered=False)
    full = pa.Table.from_batches(t.data).combine_chunks().sort_by("v")     # shards with disjoint key ranges
    ts = MemoryTable.try_new(SCHEMA, full.to_batches())
    rows = run_exchange_emulated(gpu_ctx, ts, 4, plan, key_cols=2)
    assert rows is not None
    check_rows("exchange 2 keys", rows, single, ordered=False)
