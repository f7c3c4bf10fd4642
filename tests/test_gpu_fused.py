"""The fused scan->filter->aggregate kernel (csrc/fused.cu): strategy selection + parity with the oracle on
the shapes it claims (dense private tables, HBM hash table, carry mode, reserved-code key, empty ranges,
dictionary-coded Utf8 predicates/keys), and fall-back to the generic path on the shapes it must refuse."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import qref
from qurious_b200 import tpch
from qurious_b200.datatypes import ScalarValue
from qurious_b200.physical.expr import (AvgAggregateExpr, CastExpr, Column, CountAggregateExpr, Literal, MaxAggregateExpr,
                                        MinAggregateExpr, SumAggregateExpr, avg_return_type)
from qurious_b200.physical.plan import Filter, HashAggregate, MemoryTable, NoGroupingAggregate, Scan
from tests.cases import bx, check_rows, lit, rows_of

pytestmark = pytest.mark.gpu
DEC = pa.decimal128(15, 2)


def run_both(plan, ctx, ordered=False):
    got, ref = plan.execute(ctx), qref.execute(plan)
    if got and ref:
        assert got[0].schema.types == ref[0].schema.types
    check_rows("fused", rows_of(got), rows_of(ref), ordered=ordered)
    return plan.last_strategy()


@pytest.mark.parametrize("batch_rows", [1024, 700, None])
def test_q1_q6_use_the_fused_kernel(gpu_ctx, batch_rows):
    db = tpch.generate(0.02, batch_rows=batch_rows)
    s = run_both(tpch.q6_plan(db), gpu_ctx)
    assert "fused_scan_agg[dense-private" in s, s
    s = run_both(tpch.q1_plan(db), gpu_ctx)
    assert "fused_scan_agg[dense-private" in s, s


def _table(n, seed=3, nulls=False):
    rng = np.random.default_rng(seed)
    cols = {
        "k": pa.array(rng.integers(0, 5, n), pa.int64()),
        "d": pa.array(rng.integers(9000, 9100, n).astype(np.int32), pa.int32()).cast(pa.date32()),
        "s": pa.array([["x", "BUILDING", "yy", ""][i] for i in rng.integers(0, 4, n)], pa.string()),
        "v": pa.array(rng.integers(-10**6, 10**6, n), pa.int64()),
        "p": pa.Array.from_buffers(DEC, n, [None, pa.py_buffer(np.stack([rng.integers(0, 10**7, n), np.zeros(n, np.int64)], 1).tobytes())]),
        "f": pa.array(rng.random(n), pa.float64()),
        "i32": pa.array(rng.integers(-100, 100, n).astype(np.int32), pa.int32()),
        "u16": pa.array(rng.integers(0, 60000, n).astype(np.uint16), pa.uint16()),
    }
    schema = pa.schema([(k, v.type) for k, v in cols.items()])
    full = pa.record_batch(list(cols.values()), schema=schema)
    cut = n // 3
    return MemoryTable.try_new(schema, [full.slice(0, cut), full.slice(cut, 0), full.slice(cut)])


def C(t, name):
    return Column(name, t.schema.get_field_index(name))


def test_dense_groups_strings_dates_and_minmax(gpu_ctx):
    t = _table(20000)
    pred = bx(bx(C(t, "s"), "Eq", lit("BUILDING")), "And", bx(C(t, "d"), "Lt", CastExpr(lit("1994-11-01"), pa.date32())))
    one = CastExpr(lit(1), pa.decimal128(20, 0))
    arg = bx(C(t, "p"), "Mul", bx(one, "Sub", C(t, "p")))          # p * (1 - p): negative values, Decimal128(38,4)
    aggs = [SumAggregateExpr(arg, pa.decimal128(38, 4)), MinAggregateExpr(C(t, "v"), pa.int64()),
            MaxAggregateExpr(C(t, "p"), DEC), AvgAggregateExpr(C(t, "p"), DEC, avg_return_type(DEC)),
            CountAggregateExpr(C(t, "v")), AvgAggregateExpr(C(t, "f"), pa.float64(), pa.float64()),
            SumAggregateExpr(bx(C(t, "i32"), "Mul", C(t, "i32")), pa.int32())]
    # Sum over Int32 is rejected by the reference (sum.rs:37-50): drop it, keep the rest
    aggs = aggs[:-1]
    types = [a.return_type for a in aggs]
    schema = pa.schema([("k", pa.int64()), ("s", pa.string())] + [(f"a{i}", ty) for i, ty in enumerate(types)])
    plan = HashAggregate(schema, Scan(t.schema, t, None, pred), [C(t, "k"), C(t, "s")], aggs)
    got, ref = plan.execute(gpu_ctx), qref.execute(plan)
    # 5 x 4 key domain x 7 accumulators does not fit per-thread private tables: HBM hash table
    assert "fused_scan_agg[hbm-hash" in plan.last_strategy(), plan.last_strategy()
    schema1 = pa.schema([("s", pa.string())] + [(f"a{i}", ty) for i, ty in enumerate(types)])
    plan1 = HashAggregate(schema1, Scan(t.schema, t, None, pred), [C(t, "s")], aggs)
    got1, ref1 = plan1.execute(gpu_ctx), qref.execute(plan1)
    assert "fused_scan_agg[dense-private" in plan1.last_strategy(), plan1.last_strategy()
    for a, b in zip(sorted(rows_of(got1)), sorted(rows_of(ref1))):
        assert a[:6] == b[:6] and abs(a[6] - b[6]) <= 1e-12 * abs(b[6])
    g, r = sorted(rows_of(got)), sorted(rows_of(ref))
    assert len(g) == len(r)
    for a, b in zip(g, r):
        assert a[:7] == b[:7]
        assert abs(a[7] - b[7]) <= 1e-12 * abs(b[7])
    # Filter node above the Scan is folded into the same kernel
    plan2 = HashAggregate(schema, Filter(Scan(t.schema, t, None, None), pred), [C(t, "k"), C(t, "s")], aggs)
    got2 = plan2.execute(gpu_ctx)
    assert "fused_scan_agg" in plan2.last_strategy()
    assert sorted(x[:7] for x in rows_of(got2)) == sorted(x[:7] for x in r)


def test_empty_range_and_unknown_string(gpu_ctx):
    t = _table(5000)
    for pred in (bx(C(t, "s"), "Eq", lit("nope")),
                 bx(bx(C(t, "k"), "Gt", lit(3)), "And", bx(C(t, "k"), "Lt", lit(2))),
                 bx(C(t, "v"), "Gt", lit(2**63 - 1))):
        schema = pa.schema([("c", pa.int64()), ("s", pa.int64()), ("m", pa.int64())])
        plan = NoGroupingAggregate(schema, Scan(t.schema, t, None, pred),
                                   [CountAggregateExpr(lit(1)), SumAggregateExpr(C(t, "v"), pa.int64()),
                                    MinAggregateExpr(C(t, "v"), pa.int64())])
        s = run_both(plan, gpu_ctx, ordered=True)
        assert "fused_scan_agg" in s, s
        gschema = pa.schema([("k", pa.int64()), ("c", pa.int64())])
        gplan = HashAggregate(gschema, Scan(t.schema, t, None, pred), [C(t, "k")], [CountAggregateExpr(lit(1))])
        run_both(gplan, gpu_ctx)
        # typed MIN / MAX start values when no row qualifies (min.rs / max.rs NATIVE::MAX / MIN; quirk Q4): Decimal128 keeps
        # i128::MAX / MIN, Int32 / UInt16 their own type's limits -- not the kernels' 64-bit working seed
        tschema = pa.schema([("mn_p", DEC), ("mx_p", DEC), ("mn_i", pa.int32()), ("mx_i", pa.int32()), ("mn_u", pa.uint16()),
                             ("mx_u", pa.uint16())])
        tplan = NoGroupingAggregate(tschema, Scan(t.schema, t, None, pred),
                                    [MinAggregateExpr(C(t, "p"), DEC), MaxAggregateExpr(C(t, "p"), DEC),
                                     MinAggregateExpr(C(t, "i32"), pa.int32()), MaxAggregateExpr(C(t, "i32"), pa.int32()),
                                     MinAggregateExpr(C(t, "u16"), pa.uint16()), MaxAggregateExpr(C(t, "u16"), pa.uint16())])
        s = run_both(tplan, gpu_ctx, ordered=True)
        assert "fused_scan_agg" in s, s


def test_hash_mode_full_range_keys_and_carry(gpu_ctx):
    """keys spanning the whole int64 range (one of them packs to the reserved code), values large enough
    that group totals overflow int64 (128-bit carry path)."""
    rng = np.random.default_rng(11)
    n = 200_000
    base = np.array([-2**63, 2**63 - 1, 0, -1, 12345678901234], dtype=np.int64)
    k = np.concatenate([base[rng.integers(0, 5, n // 2)], rng.integers(-2**63, 2**63 - 1, n // 2)])
    big = rng.integers(2**58, 2**59, n)
    schema = pa.schema([("k", pa.int64()), ("big", pa.decimal128(18, 0)), ("v", pa.int64())])
    bigcol = pa.Array.from_buffers(pa.decimal128(18, 0), n, [None, pa.py_buffer(np.stack([big, np.zeros(n, np.int64)], 1).tobytes())])
    t = MemoryTable.try_new(schema, [pa.record_batch([pa.array(k), bigcol, pa.array(big // 7)], schema=schema)])
    out = pa.schema([("k", pa.int64()), ("s", pa.decimal128(18, 0)), ("c", pa.int64()), ("mx", pa.int64())])
    plan = HashAggregate(out, Scan(schema, t, None, None), [Column("k", 0)],
                         [SumAggregateExpr(Column("big", 1), pa.decimal128(18, 0)), CountAggregateExpr(lit(1)),
                          MaxAggregateExpr(Column("v", 2), pa.int64())])
    got = rows_of(plan.execute(gpu_ctx))
    strat = plan.last_strategy()
    exp = {}
    for kk, b in zip(k.tolist(), big.tolist()):
        e = exp.setdefault(kk, [0, 0, -2**63])
        e[0] += b
        e[1] += 1
        e[2] = max(e[2], b // 7)
    assert len(got) == len(exp)
    for row in got:
        e = exp[row[0]]
        assert int(row[1]) == e[0] and row[2] == e[1] and row[3] == e[2], (row, e)
    assert "fused_scan_agg[hbm-hash" in strat, strat


def test_hash_table_growth(gpu_ctx):
    rng = np.random.default_rng(2)
    n = 6_000_000
    k = rng.integers(0, 3_000_000, n) * 1_000_003
    v = rng.integers(-1000, 1000, n)
    schema = pa.schema([("k", pa.int64()), ("v", pa.int64())])
    t = MemoryTable.try_new(schema, [pa.record_batch([pa.array(k), pa.array(v)], schema=schema)])
    out = pa.schema([("k", pa.int64()), ("s", pa.int64()), ("c", pa.int64())])
    plan = HashAggregate(out, Scan(schema, t, None, bx(Column("v", 1), "GtEq", lit(-900))), [Column("k", 0)],
                         [SumAggregateExpr(Column("v", 1), pa.int64()), CountAggregateExpr(Column("v", 1))])
    tab = plan.execute(gpu_ctx)
    assert "fused_scan_agg[hbm-hash" in plan.last_strategy(), plan.last_strategy()
    gk = np.concatenate([b.column(0).to_numpy() for b in tab])
    gs = np.concatenate([b.column(1).to_numpy() for b in tab])
    gc = np.concatenate([b.column(2).to_numpy() for b in tab])
    m = v >= -900
    uk, inv = np.unique(k[m], return_inverse=True)
    es = np.bincount(inv, weights=None, minlength=len(uk))
    ssum = np.zeros(len(uk), dtype=np.int64)
    np.add.at(ssum, inv, v[m])
    order = np.argsort(gk)
    assert np.array_equal(gk[order], uk)
    assert np.array_equal(gs[order], ssum)
    assert np.array_equal(gc[order], es)


def test_shapes_the_fused_kernel_must_refuse(gpu_ctx):
    t = _table(3000)
    src = Scan(t.schema, t, None, bx(C(t, "k"), "NotEq", lit(2)))          # != is not a range
    plan = HashAggregate(pa.schema([("k", pa.int64()), ("c", pa.int64())]), src, [C(t, "k")], [CountAggregateExpr(lit(1))])
    s = run_both(plan, gpu_ctx)
    assert "generic" in s, s
    src = Scan(t.schema, t, None, bx(bx(C(t, "k"), "Eq", lit(2)), "Or", bx(C(t, "k"), "Eq", lit(3))))
    plan = NoGroupingAggregate(pa.schema([("c", pa.int64())]), src, [CountAggregateExpr(lit(1))])
    s = run_both(plan, gpu_ctx)
    assert "generic" in s, s
    # computed group key
    plan = HashAggregate(pa.schema([("k", pa.int64()), ("c", pa.int64())]), Scan(t.schema, t, None, None),
                         [bx(C(t, "k"), "Add", lit(1))], [CountAggregateExpr(lit(1))])
    s = run_both(plan, gpu_ctx)
    assert "generic" in s, s


@pytest.mark.parametrize("batch_rows", [1024, None])
def test_q3_uses_the_fused_join_probe_aggregate(gpu_ctx, batch_rows):
    db = tpch.generate(0.02, batch_rows=batch_rows)
    plan = tpch.q3_plan(db)
    got, ref = plan.execute(gpu_ctx), qref.execute(plan)
    assert "fused_join_probe_agg[unique-build" in plan.last_strategy(), plan.last_strategy()
    assert got[0].schema.types == ref[0].schema.types
    # same rows AND the same (first-occurrence) order as the generic operators / the oracle
    check_rows("q3", rows_of(got), rows_of(ref), ordered=True)


def test_join_aggregate_falls_back_on_duplicate_build_keys(gpu_ctx):
    from qurious_b200.datatypes import JoinType
    from qurious_b200.physical.plan import HashJoinExec
    rng = np.random.default_rng(4)
    nb, n_probe = 500, 20000
    bs = pa.schema([("bk", pa.int64()), ("tag", pa.int64())])
    ps = pa.schema([("pk", pa.int64()), ("v", pa.int64())])
    for dup in (False, True):
        bk = np.arange(nb) * 3 if not dup else rng.integers(0, 200, nb) * 3
        build = MemoryTable.try_new(bs, [pa.record_batch([pa.array(bk), pa.array(rng.integers(0, 9, nb))], schema=bs)])
        probe = MemoryTable.try_new(ps, [pa.record_batch([pa.array(rng.integers(0, 3 * nb, n_probe)),
                                                          pa.array(rng.integers(-1000, 1000, n_probe))], schema=ps)])
        j = HashJoinExec.try_new(Scan(bs, build, None, None), Scan(ps, probe, None, bx(Column("v", 1), "GtEq", lit(-500))),
                                 JoinType.Inner, [(Column("bk", 0), Column("pk", 0))], None)
        out = pa.schema([("pk", pa.int64()), ("tag", pa.int64()), ("s", pa.int64()), ("c", pa.int64()), ("mn", pa.int64())])
        keys = [Column("pk", 2), Column("tag", 1)] if not dup else [Column("pk", 2)]
        if dup:
            out = pa.schema([("pk", pa.int64()), ("s", pa.int64()), ("c", pa.int64()), ("mn", pa.int64())])
        plan = HashAggregate(out, j, keys, [SumAggregateExpr(Column("v", 3), pa.int64()), CountAggregateExpr(lit(1)),
                                            MinAggregateExpr(Column("v", 3), pa.int64())])
        s = run_both(plan, gpu_ctx, ordered=True)
        assert ("fused_join_probe_agg" in s) == (not dup), s


# ---- packed accumulators (FParams::pack_mask): SUM(l_quantity) and SUM(l_discount) of the Q1 shape ride in the row-count
# ---- word of the private tables only when the column statistics prove it safe; every variant must equal the oracle
def _q1_db(mutate):
    t = tpch.gen_lineitem(0.02, columns=tpch.Q1_COLUMNS)
    mutate(t.cols)
    li = MemoryTable.try_new(t.schema, tpch.to_arrow(t, 1024))
    return tpch.Database(0.02, None, None, li)


def test_q1_packs_small_sums_into_the_count_word(gpu_ctx):
    s = run_both(tpch.q1_plan(_q1_db(lambda c: None)), gpu_ctx)
    assert "2 accs packed into the count word" in s, s


def test_q1_negative_quantity_is_not_packed(gpu_ctx):
    def neg(c):
        c["l_quantity"][::97] = -c["l_quantity"][::97]
    s = run_both(tpch.q1_plan(_q1_db(neg)), gpu_ctx)
    assert "shape-specialised" in s and "packed" not in s, s


def test_q1_wide_discount_is_not_packed(gpu_ctx):
    def wide(c):
        c["l_extendedprice"][:] = 1                   # keeps price * (100 - disc) * (100 + tax) inside int64
        c["l_tax"][:] = 0
        c["l_discount"][::5] = 10 ** 14 + 7          # a per-thread partial sum no longer fits a 48-bit field
    s = run_both(tpch.q1_plan(_q1_db(wide)), gpu_ctx)
    assert "fused_scan_agg[dense-private/shape-specialised" in s and "packed" not in s, s


def test_q1_packed_fields_at_their_bounds(gpu_ctx):
    def big(c):
        c["l_quantity"][:] = 99_999_999_00          # every row at the column maximum: the bound the host sizes the field by
        c["l_discount"][:] = 10
    s = run_both(tpch.q1_plan(_q1_db(big)), gpu_ctx)
    assert "fused_scan_agg[dense-private" in s, s


def test_small_string_keys_one_launch_gather(gpu_ctx):
    """<= 1024 result groups with Utf8 keys take the single-launch gather (k_str_take_small), more take the
    lens -> scan -> copy path; both must give the reference's strings (including empty ones)."""
    rng = np.random.default_rng(11)
    for n_keys in (3, 1024, 1500):
        n = 6000
        words = ["", "a", "BUILDING", "x" * 37] + [f"k{i:05d}" for i in range(n_keys - 4)] if n_keys > 4 else ["", "a", "x" * 37]
        ks = pa.array([words[i] for i in rng.integers(0, len(words), n)], pa.string())
        vs = pa.array(rng.integers(-1000, 1000, n), pa.int64())
        schema = pa.schema([("s", pa.string()), ("v", pa.int64())])
        t = MemoryTable.try_new(schema, [pa.record_batch([ks, vs], schema=schema)])
        out = pa.schema([("s", pa.string()), ("sum_v", pa.int64()), ("n", pa.int64())])
        plan = HashAggregate(out, Scan(schema, t, None, None), [Column("s", 0)],
                             [SumAggregateExpr(Column("v", 1), pa.int64()), CountAggregateExpr(Column("v", 1))])
        run_both(plan, gpu_ctx)


def test_unregistered_dense_shape_is_specialised_at_run_time(gpu_ctx, monkeypatch):
    """A DENSE plan whose shape has no ahead-of-time kernel (MIN / MAX / decimal and Float64 sums over a string key) runs
    SpecBody instantiated for its signature by NVRTC (csrc/fused_jit.cu); QGPU_JIT=0 runs the interpreted body instead.
    Both equal the oracle."""
    t = _table(30000, seed=11)
    pred = bx(C(t, "d"), "Lt", CastExpr(lit("1994-11-01"), pa.date32()))
    aggs = [SumAggregateExpr(C(t, "p"), DEC), MinAggregateExpr(C(t, "v"), pa.int64()), MaxAggregateExpr(C(t, "p"), DEC),
            CountAggregateExpr(C(t, "v")), AvgAggregateExpr(C(t, "f"), pa.float64(), pa.float64())]
    schema = pa.schema([("s", pa.string())] + [(f"a{i}", a.return_type) for i, a in enumerate(aggs)])

    def make():
        return HashAggregate(schema, Scan(t.schema, t, None, pred), [C(t, "s")], aggs)
    ref = sorted(rows_of(qref.execute(make())))

    def check(plan):
        got = sorted(rows_of(plan.execute(gpu_ctx)))
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            assert a[:5] == b[:5] and abs(a[5] - b[5]) <= 1e-12 * abs(b[5])
        return plan.last_strategy()
    s = check(make())
    assert "dense-private/shape-specialised at run time" in s, s
    monkeypatch.setenv("QGPU_JIT", "0")
    s = check(make())
    assert "dense-private, " in s and "specialised" not in s, s
