"""GPU operators vs the oracle on seeded random inputs: every type, NULLs, ragged batches, all join
types.  Bit-exact for integer / decimal / boolean / string / index work; Float64 SUM/AVG within
1e-12 relative (BASELINE.json north_star: only the reduction order differs)."""
import decimal
import math

import numpy as np
import pyarrow as pa
import pytest

from oracle import qref
from qurious_b200 import QuriousError, tpch
from qurious_b200.datatypes import JoinSide, JoinType, Operator, ScalarValue
from qurious_b200.physical.expr import (AvgAggregateExpr, BinaryExpr, CaseExpr, CastExpr, Column, CountAggregateExpr,
                                        IsNotNull, IsNull, Literal, MaxAggregateExpr, MinAggregateExpr, Negative,
                                        SumAggregateExpr, avg_return_type)
from qurious_b200.physical.plan import (Filter, HashAggregate, HashJoinExec, JoinFilter, MemoryTable,
                                        NoGroupingAggregate, Projection, Scan)
from tests.cases import bx, check_rows, lit, rows_of, sort_key

pytestmark = pytest.mark.gpu
FLOAT_RTOL = 1e-12  # BASELINE.json north_star tolerance for Float64 SUM/AVG


def dec_array(raw, p, s, valid):
    vals = [decimal.Decimal(int(v)).scaleb(-s) if ok else None for v, ok in zip(raw, valid)]
    return pa.array(vals, type=pa.decimal128(p, s))


def random_table(rng, n, null_frac=0.2, splits=None, nullable=True):
    def mask():
        return rng.random(n) >= (null_frac if nullable else 0.0)

    def arr(vals, t):
        m = mask()
        return pa.array([v if ok else None for v, ok in zip(vals.tolist(), m)], type=t)
    cols = {
        "i8": arr(rng.integers(-128, 128, n), pa.int8()),
        "i16": arr(rng.integers(-32768, 32768, n), pa.int16()),
        "i32": arr(rng.integers(-50, 50, n), pa.int32()),
        "i64": arr(rng.integers(-2**62, 2**62, n), pa.int64()),
        "k64": arr(rng.integers(0, 12, n), pa.int64()),
        "u8": arr(rng.integers(0, 256, n), pa.uint8()),
        "u32": arr(rng.integers(0, 2**32, n), pa.uint32()),
        "u64": arr(rng.integers(0, 2**63, n).astype(np.uint64) * 2, pa.uint64()),
        "f32": arr(rng.normal(0, 100, n).astype(np.float32), pa.float32()),
        "f64": arr(rng.normal(0, 1e6, n), pa.float64()),
        "b": arr(rng.integers(0, 2, n).astype(bool), pa.bool_()),
        "b2": arr(rng.integers(0, 2, n).astype(bool), pa.bool_()),
        "d32": arr(rng.integers(8000, 11000, n), pa.int32()).cast(pa.date32()),
        "dec": dec_array(rng.integers(-10**12, 10**12, n), 15, 2, mask()),
        "dec2": dec_array(rng.integers(0, 11, n), 15, 2, mask()),
        "wide": dec_array([int(x) * 10**18 + int(y) for x, y in zip(rng.integers(-10**15, 10**15, n), rng.integers(0, 10**18, n))],
                          38, 4, mask()),
        "s": pa.array([None if not ok else ["", "a", "ab", "BUILDING", "b", "zz top", "été"][v]
                       for v, ok in zip(rng.integers(0, 7, n), mask())], type=pa.string()),
    }
    schema = pa.schema([pa.field(k, v.type, nullable) for k, v in cols.items()])
    full = pa.record_batch(list(cols.values()), schema=schema)
    if splits is None:
        splits = [n]
    batches, off = [], 0
    for s_ in splits:
        batches.append(full.slice(off, s_))
        off += s_
    return MemoryTable.try_new(schema, batches)


def c(t, name):
    return Column(name, t.schema.get_field_index(name))


def assert_same(name, plan, ctx, ordered=True, float_cols=()):
    got = plan.execute(ctx)
    ref = qref.execute(plan)
    g, r = rows_of(got), rows_of(ref)
    assert len(g) == len(r), f"{name}: {len(g)} rows vs {len(r)}"
    if got and ref:
        assert got[0].schema.types == ref[0].schema.types, f"{name}: {got[0].schema.types} vs {ref[0].schema.types}"
    if not ordered:
        g, r = sorted(g, key=sort_key), sorted(r, key=sort_key)
    for i, (a, b) in enumerate(zip(g, r)):
        for j, (x, y) in enumerate(zip(a, b)):
            if isinstance(y, float) and x is not None and y is not None:
                if math.isnan(y):
                    assert math.isnan(x), f"{name} row {i} col {j}: {x} vs {y}"
                elif j in float_cols:
                    assert abs(x - y) <= FLOAT_RTOL * max(abs(y), 1e-300), f"{name} row {i} col {j}: {x} vs {y}"
                else:
                    assert x == y or (x != x and y != y), f"{name} row {i} col {j}: {x!r} vs {y!r}"
            else:
                assert x == y, f"{name} row {i} col {j}: {x!r} vs {y!r}\n{a}\n{b}"


EXPRS = [
    # (name, builder(t), result type)
    ("i32<lit", lambda t: bx(c(t, "i32"), "Lt", Literal(ScalarValue.Int32(3))), pa.bool_()),
    ("i64>=i64", lambda t: bx(c(t, "i64"), "GtEq", c(t, "i64")), pa.bool_()),
    ("u64>lit", lambda t: bx(c(t, "u64"), "Gt", Literal(ScalarValue.UInt64(2**63 + 5))), pa.bool_()),
    ("f64 total order", lambda t: bx(c(t, "f64"), "LtEq", lit(0.0)), pa.bool_()),
    ("f32==f32", lambda t: bx(c(t, "f32"), "Eq", c(t, "f32")), pa.bool_()),
    ("dec<dec2", lambda t: bx(c(t, "dec"), "Lt", c(t, "dec2")), pa.bool_()),
    ("wide>=wide", lambda t: bx(c(t, "wide"), "GtEq", CastExpr(lit(0), pa.decimal128(38, 4))), pa.bool_()),
    ("s=BUILDING", lambda t: bx(c(t, "s"), "Eq", lit("BUILDING")), pa.bool_()),
    ("s<lit", lambda t: bx(c(t, "s"), "Lt", lit("b")), pa.bool_()),
    ("s!=s", lambda t: bx(c(t, "s"), "NotEq", c(t, "s")), pa.bool_()),
    ("date<cast", lambda t: bx(c(t, "d32"), "Lt", CastExpr(lit("1995-03-15"), pa.date32())), pa.bool_()),
    ("b and b2", lambda t: bx(c(t, "b"), "And", c(t, "b2")), pa.bool_()),
    ("b or b2", lambda t: bx(c(t, "b"), "Or", c(t, "b2")), pa.bool_()),
    ("b = b2", lambda t: bx(c(t, "b"), "Eq", c(t, "b2")), pa.bool_()),
    ("i8+i8 wraps", lambda t: bx(c(t, "i8"), "Add", c(t, "i8")), pa.int8()),
    ("i16*i16 wraps", lambda t: bx(c(t, "i16"), "Mul", c(t, "i16")), pa.int16()),
    ("i64*i64 wraps", lambda t: bx(c(t, "i64"), "Mul", c(t, "i64")), pa.int64()),
    ("u8-u8 wraps", lambda t: bx(c(t, "u8"), "Sub", c(t, "u8")), pa.uint8()),
    ("u32*u32 wraps", lambda t: bx(c(t, "u32"), "Mul", c(t, "u32")), pa.uint32()),
    ("i64/lit", lambda t: bx(c(t, "i64"), "Div", lit(-7)), pa.int64()),
    ("i64%lit", lambda t: bx(c(t, "i64"), "Mod", lit(-7)), pa.int64()),
    ("f64/f64", lambda t: bx(c(t, "f64"), "Div", c(t, "f64")), pa.float64()),
    ("f64*f64", lambda t: bx(c(t, "f64"), "Mul", lit(1.5)), pa.float64()),
    ("f32+f32", lambda t: bx(c(t, "f32"), "Add", c(t, "f32")), pa.float32()),
    ("f64%lit", lambda t: bx(c(t, "f64"), "Mod", lit(3.25)), pa.float64()),
    ("dec+dec2", lambda t: bx(c(t, "dec"), "Add", c(t, "dec2")), pa.decimal128(16, 2)),
    ("dec*(1-dec2)", lambda t: bx(c(t, "dec"), "Mul", bx(CastExpr(lit(1), pa.decimal128(20, 0)), "Sub", c(t, "dec2"))),
     pa.decimal128(38, 4)),
    ("wide*dec2 wraps", lambda t: bx(c(t, "wide"), "Mul", c(t, "dec2")), pa.decimal128(38, 6)),
    ("wide-dec rescale", lambda t: bx(c(t, "wide"), "Sub", c(t, "dec")), pa.decimal128(38, 4)),
    ("dec/dec2 -> f64", lambda t: bx(c(t, "dec"), "Div", bx(c(t, "dec2"), "Add", CastExpr(lit(1), pa.decimal128(15, 2)))),
     pa.float64()),
    ("dec%lit", lambda t: bx(c(t, "dec"), "Mod", CastExpr(lit(7), pa.decimal128(15, 2))), pa.decimal128(15, 2)),
    ("neg i32", lambda t: Negative(c(t, "i32")), pa.int32()),
    ("neg dec", lambda t: Negative(c(t, "wide")), pa.decimal128(38, 4)),
    ("neg f64", lambda t: Negative(c(t, "f64")), pa.float64()),
    ("isnull", lambda t: IsNull(c(t, "s")), pa.bool_()),
    ("isnotnull", lambda t: IsNotNull(c(t, "wide")), pa.bool_()),
    ("cast i32->i64", lambda t: CastExpr(c(t, "i32"), pa.int64()), pa.int64()),
    ("cast i32->dec", lambda t: CastExpr(c(t, "i32"), pa.decimal128(15, 2)), pa.decimal128(15, 2)),
    ("cast i32->f64", lambda t: CastExpr(c(t, "i32"), pa.float64()), pa.float64()),
    ("cast dec->f64", lambda t: CastExpr(c(t, "dec"), pa.float64()), pa.float64()),
    ("cast dec->dec(20,1) rounds", lambda t: CastExpr(c(t, "dec"), pa.decimal128(20, 1)), pa.decimal128(20, 1)),
    ("cast dec->i64 truncates", lambda t: CastExpr(c(t, "dec"), pa.int64()), pa.int64()),
    ("cast f32->dec", lambda t: CastExpr(c(t, "f32"), pa.decimal128(15, 2)), pa.decimal128(15, 2)),
    ("cast f64->i64", lambda t: CastExpr(c(t, "f64"), pa.int64()), pa.int64()),
    ("case", lambda t: CaseExpr([(bx(c(t, "i32"), "Gt", Literal(ScalarValue.Int32(10))), c(t, "i64")),
                                 (c(t, "b"), lit(7))], Negative(c(t, "i64"))), pa.int64()),
    ("kleene tree", lambda t: bx(bx(c(t, "b"), "And", bx(c(t, "i32"), "Gt", Literal(ScalarValue.Int32(0)))), "Or",
                                 bx(IsNull(c(t, "dec")), "And", c(t, "b2"))), pa.bool_()),
    ("const folded", lambda t: bx(bx(lit(2), "Mul", lit(21)), "Add", CastExpr(c(t, "i32"), pa.int64())), pa.int64()),
]


@pytest.mark.parametrize("nullable", [True, False])
def test_expressions_random(gpu_ctx, nullable):
    rng = np.random.default_rng(7 if nullable else 8)
    t = random_table(rng, 777, splits=[300, 1, 0, 476], nullable=nullable)
    schema = pa.schema([pa.field(f"e{i}", ty, True) for i, (_, _, ty) in enumerate(EXPRS)])
    # one projection per expression first (so that a failure names the expression) ...
    for i, (name, build, ty) in enumerate(EXPRS):
        plan = Projection(pa.schema([pa.field("r", ty, True)]), Scan(t.schema, t, None, None), [build(t)])
        assert_same(f"expr[{name}]", plan, gpu_ctx)
    # ... then all of them in a single Projection
    plan = Projection(schema, Scan(t.schema, t, None, None), [b(t) for _, b, _ in EXPRS])
    assert_same("all expressions", plan, gpu_ctx)


def test_filter_random_all_columns_survive(gpu_ctx):
    rng = np.random.default_rng(11)
    t = random_table(rng, 5000, splits=[1024, 1024, 1024, 1024, 904])
    pred = bx(bx(c(t, "i32"), "Gt", Literal(ScalarValue.Int32(-10))), "And",
              bx(bx(c(t, "s"), "Eq", lit("BUILDING")), "Or", bx(c(t, "dec2"), "LtEq", CastExpr(lit(0.07), pa.decimal128(15, 2)))))
    assert_same("scan filter", Scan(t.schema, t, None, pred), gpu_ctx)
    assert_same("Filter(Filter)", Filter(Filter(Scan(t.schema, t, None, None), pred), IsNotNull(c(t, "wide"))), gpu_ctx)
    assert_same("projection pushdown arg", Scan(pa.schema([t.schema.field("s"), t.schema.field("i64")]), t, ["s", "i64"],
                                              IsNotNull(Column("s", 0))), gpu_ctx)
    # nothing survives / everything survives / empty table
    assert_same("none", Scan(t.schema, t, None, bx(c(t, "i32"), "Gt", Literal(ScalarValue.Int32(1000)))), gpu_ctx)
    assert_same("all", Scan(t.schema, t, None, bx(IsNull(c(t, "i8")), "Or", IsNotNull(c(t, "i8")))), gpu_ctx)


def agg_suite(t):
    avg_d = avg_return_type(pa.decimal128(15, 2))
    aggs = [
        SumAggregateExpr(c(t, "k64"), pa.int64()), SumAggregateExpr(c(t, "i64"), pa.int64()),
        SumAggregateExpr(c(t, "u64"), pa.uint64()), SumAggregateExpr(c(t, "f64"), pa.float64()),
        SumAggregateExpr(c(t, "dec"), pa.decimal128(15, 2)), SumAggregateExpr(c(t, "wide"), pa.decimal128(38, 4)),
        SumAggregateExpr(bx(c(t, "dec"), "Mul", c(t, "dec2")), pa.decimal128(31, 4)),
        CountAggregateExpr(c(t, "s")), CountAggregateExpr(lit(1)),
        AvgAggregateExpr(c(t, "dec"), pa.decimal128(15, 2), avg_d), AvgAggregateExpr(c(t, "f64"), pa.float64(), pa.float64()),
        MinAggregateExpr(c(t, "i64"), pa.int64()), MaxAggregateExpr(c(t, "i64"), pa.int64()),
        MinAggregateExpr(c(t, "i8"), pa.int8()), MaxAggregateExpr(c(t, "u32"), pa.uint32()),
        MinAggregateExpr(c(t, "u64"), pa.uint64()), MaxAggregateExpr(c(t, "f64"), pa.float64()),
        MinAggregateExpr(c(t, "f32"), pa.float32()),
        MinAggregateExpr(c(t, "wide"), pa.decimal128(38, 4)), MaxAggregateExpr(c(t, "wide"), pa.decimal128(38, 4)),
        MinAggregateExpr(c(t, "dec"), pa.decimal128(15, 2)),
    ]
    types = [a.return_type for a in aggs]
    float_cols = [i for i, ty in enumerate(types) if ty == pa.float64() and not isinstance(aggs[i], (MinAggregateExpr, MaxAggregateExpr))]
    return aggs, types, float_cols


@pytest.mark.parametrize("nullable", [True, False])
def test_aggregate_random(gpu_ctx, nullable):
    rng = np.random.default_rng(21)
    t = random_table(rng, 4000, splits=[1000, 2000, 0, 1000], nullable=nullable)
    aggs, types, float_cols = agg_suite(t)
    names = [f"a{i}" for i in range(len(aggs))]
    src = Scan(t.schema, t, None, None)
    schema = pa.schema(list(zip(names, types)))
    assert_same("no grouping", NoGroupingAggregate(schema, src, aggs), gpu_ctx, float_cols=float_cols)
    filt = Scan(t.schema, t, None, bx(c(t, "i32"), "Gt", Literal(ScalarValue.Int32(1000))))
    # every row filtered: COUNT=0, SUM=NULL, MIN/MAX = sentinel, AVG NULL
    assert_same("no grouping, no rows", NoGroupingAggregate(schema, filt, aggs), gpu_ctx, float_cols=float_cols)
    for keys, key_types in ((["k64"], [pa.int64()]), (["s", "i32"], [pa.string(), pa.int32()]),
                            (["d32", "dec2", "u8"], [pa.date32(), pa.decimal128(15, 2), pa.uint8()])):
        gschema = pa.schema(list(zip(keys, key_types)) + list(zip(names, types)))
        plan = HashAggregate(gschema, src, [c(t, k) for k in keys], aggs)
        assert_same(f"group by {keys}", plan, gpu_ctx, ordered=False, float_cols=[len(keys) + i for i in float_cols])
    # first-occurrence output order is OUR contract (reference order is unspecified): oracle does the same
    plan = HashAggregate(pa.schema([("k64", pa.int64()), ("c", pa.int64())]), src, [c(t, "k64")], [CountAggregateExpr(lit(1))])
    assert_same("group order", plan, gpu_ctx, ordered=True)
    # expression keys, grouped over zero rows (one 0-row output batch)
    plan = HashAggregate(pa.schema([("k", pa.int64()), ("c", pa.int64())]), filt, [bx(c(t, "k64"), "Add", lit(1))],
                         [CountAggregateExpr(lit(1))])
    assert_same("group over zero rows", plan, gpu_ctx)


def test_aggregate_many_groups(gpu_ctx):
    """high-cardinality: forces the HBM table to grow (64 Ki -> 1 Mi slots)."""
    rng = np.random.default_rng(5)
    n = 300_000
    k = rng.integers(0, 120_000, n) * 7919 - 10**9
    v = rng.integers(-10**6, 10**6, n)
    f = rng.random(n)
    schema = pa.schema([("k", pa.int64()), ("v", pa.int64()), ("f", pa.float64())])
    t = MemoryTable.try_new(schema, [pa.record_batch([pa.array(k), pa.array(v), pa.array(f)], schema=schema)])
    K, V, F = Column("k", 0), Column("v", 1), Column("f", 2)
    out = pa.schema([("k", pa.int64()), ("s", pa.int64()), ("c", pa.int64()), ("mn", pa.int64()), ("mx", pa.int64()),
                     ("af", pa.float64())])
    plan = HashAggregate(out, Scan(schema, t, None, None), [K],
                         [SumAggregateExpr(V, pa.int64()), CountAggregateExpr(V), MinAggregateExpr(V, pa.int64()),
                          MaxAggregateExpr(V, pa.int64()), AvgAggregateExpr(F, pa.float64(), pa.float64())])
    got = rows_of(plan.execute(gpu_ctx))
    # numpy restatement (the Python-loop oracle would take minutes at this size)
    order = np.argsort(k, kind="stable")
    ks, vs, fs = k[order], v[order], f[order]
    uniq, start = np.unique(ks, return_index=True)
    end = np.append(start[1:], n)
    exp = {int(u): (int(vs[a:b].sum()), int(b - a), int(vs[a:b].min()), int(vs[a:b].max()), float(fs[a:b].sum() / (b - a)))
           for u, a, b in zip(uniq, start, end)}
    assert len(got) == len(exp)
    for row in got:
        e = exp[row[0]]
        assert row[1:5] == e[:4]
        assert abs(row[5] - e[4]) <= FLOAT_RTOL * abs(e[4])


JOIN_TYPES = [JoinType.Inner, JoinType.Left, JoinType.Right, JoinType.Full, JoinType.LeftSemi, JoinType.LeftAnti]


@pytest.mark.parametrize("join_type", JOIN_TYPES, ids=[j.name for j in JOIN_TYPES])
@pytest.mark.parametrize("with_filter", [False, True])
def test_join_random(gpu_ctx, join_type, with_filter):
    rng = np.random.default_rng(31)
    lt = random_table(rng, 600, splits=[100, 500])
    rt = random_table(rng, 900, splits=[256, 0, 300, 344])
    on = [(c(lt, "k64"), c(rt, "k64")), (c(lt, "s"), c(rt, "s"))]
    jf = None
    if with_filter:
        jf = JoinFilter(bx(Column("i32", 0), "Lt", Column("i32", 1)),
                        pa.schema([lt.schema.field("i32"), rt.schema.field("i32")]),
                        [(lt.schema.get_field_index("i32"), JoinSide.Left), (rt.schema.get_field_index("i32"), JoinSide.Right)])
    plan = HashJoinExec.try_new(Scan(lt.schema, lt, None, None), Scan(rt.schema, rt, None, None), join_type, on, jf)
    assert_same(f"join {join_type.name}", plan, gpu_ctx, ordered=True)
    # with pushed-down filters on both sides and a projection + aggregate on top (late materialisation path)
    l_scan = Scan(lt.schema, lt, None, IsNotNull(c(lt, "dec")))
    r_scan = Scan(rt.schema, rt, None, bx(c(rt, "i32"), "Gt", Literal(ScalarValue.Int32(-20))))
    j = HashJoinExec.try_new(l_scan, r_scan, join_type, [(c(lt, "k64"), c(rt, "k64"))], jf)
    agg = HashAggregate(pa.schema([("k", pa.int64()), ("n", pa.int64()), ("sd", pa.decimal128(15, 2))]), j,
                        [Column("k64", lt.schema.get_field_index("k64"))],  # left-side columns come first in the join schema
                        [CountAggregateExpr(lit(1)),
                         SumAggregateExpr(Column("dec", lt.schema.get_field_index("dec")), pa.decimal128(15, 2))])
    assert_same(f"join {join_type.name} -> aggregate", agg, gpu_ctx, ordered=False)


def test_join_heavy_duplicates_order(gpu_ctx):
    """many build rows per key: chains must come out ascending by build row (hash_join.rs:474-512)."""
    n = 3000
    lk = (np.arange(n) * 7) % 5
    rk = np.arange(40) % 7
    ls = pa.schema([("k", pa.int64()), ("row", pa.int64())])
    rs = pa.schema([("k", pa.int64()), ("row", pa.int64())])
    lt = MemoryTable.try_new(ls, [pa.record_batch([pa.array(lk), pa.array(np.arange(n))], schema=ls)])
    rt = MemoryTable.try_new(rs, [pa.record_batch([pa.array(rk), pa.array(np.arange(40))], schema=rs)])
    plan = HashJoinExec.try_new(Scan(ls, lt, None, None), Scan(rs, rt, None, None), JoinType.Inner,
                                [(Column("k", 0), Column("k", 0))], None)
    assert_same("dup order", plan, gpu_ctx, ordered=True)


def expect_error(plan, ctx, kind):
    with pytest.raises(QuriousError) as gi:
        plan.execute(ctx)
    assert gi.value.kind == kind, gi.value
    with pytest.raises(qref.QError) as oi:
        qref.execute(plan)
    assert oi.value.kind == kind, oi.value


def test_error_behaviour(gpu_ctx):
    rng = np.random.default_rng(3)
    t = random_table(rng, 50, nullable=False)
    src = Scan(t.schema, t, None, None)
    one = lambda e, ty: Projection(pa.schema([("r", ty)]), src, [e])  # noqa: E731
    expect_error(one(bx(c(t, "i64"), "Div", lit(0)), pa.int64()), gpu_ctx, "ArrowError")          # divide by zero
    expect_error(one(bx(c(t, "i32"), "Eq", c(t, "i64")), pa.bool_()), gpu_ctx, "ArrowError")      # type mismatch
    expect_error(one(bx(c(t, "dec"), "Eq", c(t, "wide")), pa.bool_()), gpu_ctx, "ArrowError")     # (p,s) mismatch
    expect_error(one(bx(c(t, "i32"), "Add", c(t, "i64")), pa.int64()), gpu_ctx, "ArrowError")
    expect_error(one(CastExpr(c(t, "i64"), pa.int8()), pa.int8()), gpu_ctx, "ArrowError")         # safe:false overflow
    expect_error(one(CastExpr(lit("not a date"), pa.date32()), pa.date32()), gpu_ctx, "ArrowError")
    expect_error(one(c(t, "i64"), pa.int32()), gpu_ctx, "ArrowError")                             # schema mismatch
    expect_error(one(Column("nope", 99), pa.int64()), gpu_ctx, "InternalError")                   # column.rs:26-31
    expect_error(HashAggregate(pa.schema([("f64", pa.float64()), ("c", pa.int64())]), src, [c(t, "f64")],
                               [CountAggregateExpr(lit(1))]), gpu_ctx, "InternalError")           # array.rs:205
    expect_error(NoGroupingAggregate(pa.schema([("s", pa.int32())]), src, [SumAggregateExpr(c(t, "i32"), pa.int32())]),
                 gpu_ctx, "InternalError")                                                        # sum.rs:47-49
    expect_error(NoGroupingAggregate(pa.schema([("a", pa.float64())]), src,
                                     [AvgAggregateExpr(c(t, "i64"), pa.int64(), pa.float64())]), gpu_ctx, "InternalError")  # avg.rs:70
    with pytest.raises(QuriousError):
        HashJoinExec.try_new(src, src, JoinType.Inner, [], None)                                  # hash_join.rs:131-133


@pytest.mark.parametrize("batch_rows", [1024, None])
def test_tpch_small_vs_oracle(gpu_ctx, batch_rows):
    db = tpch.generate(0.01, batch_rows=batch_rows)
    for q in ("q6", "q1", "q3"):
        plan = getattr(tpch, q + "_plan")(db)
        got, ref = plan.execute(gpu_ctx), qref.execute(plan)
        assert got[0].schema.types == ref[0].schema.types
        check_rows(q, rows_of(got), rows_of(ref), ordered=False)
