"""NestedLoopJoinExec (qurious/src/physical/plan/join/nest_loop_join.rs:79-300).

CPU: the oracle restatement against the reference's own two unit tests (nest_loop_join.rs:356-416) and hand-derived
vectors of the ordering rules.  GPU: parity of the library operator with the oracle through the C ABI -- all six join
types, with / without a JoinFilter, NULLs in the filter columns, empty sides, chunked cross products."""
import os

import numpy as np
import pyarrow as pa
import pytest

from oracle import qref
from qurious_b200.datatypes import JoinSide, JoinType, ScalarValue
from qurious_b200.physical.expr import CastExpr, Column, CountAggregateExpr, Literal, SumAggregateExpr
from qurious_b200.physical.plan import HashAggregate, JoinFilter, MemoryTable, NestedLoopJoinExec, Scan
from tests.cases import bx, check_rows, lit, rows_of

JOIN_TYPES = [JoinType.Inner, JoinType.Left, JoinType.Right, JoinType.Full, JoinType.LeftSemi, JoinType.LeftAnti]


def i32_table(cols):
    """build_table_scan_i32 of the reference's test_utils: Int32 columns, one batch."""
    schema = pa.schema([(n, pa.int32()) for n, _ in cols])
    batch = pa.record_batch([pa.array(v, pa.int32()) for _, v in cols], schema=schema)
    t = MemoryTable.try_new(schema, [batch])
    return Scan(schema, t, None, None)


def k_filter(op="Eq"):
    # intermediate schema (k1, k2); column_indices = [(1, Left), (0, Right)]   nest_loop_join.rs:386-402
    schema = pa.schema([pa.field("k1", pa.int32(), False), pa.field("k2", pa.int32(), False)])
    return JoinFilter(bx(Column("k1", 0), op, Column("k2", 1)), schema, [(1, JoinSide.Left), (0, JoinSide.Right)])


# ---- reference unit vectors -------------------------------------------------------------------------------
def test_oracle_left_anti_empty_right_returns_all_left():
    """nest_loop_join.rs:356-376"""
    plan = NestedLoopJoinExec.try_new(i32_table([("a1", [1, 2, 3]), ("k1", [10, 20, 30])]), i32_table([("k2", []), ("b2", [])]),
                                      JoinType.LeftAnti, None)
    out = qref.execute(plan)
    assert out[0].schema.names == ["a1", "k1"]
    assert rows_of(out) == [(1, 10), (2, 20), (3, 30)]


def test_oracle_left_semi_distinct_left_rows():
    """nest_loop_join.rs:378-416: two matching right rows must not duplicate the left row"""
    plan = NestedLoopJoinExec.try_new(i32_table([("a1", [1, 2, 3]), ("k1", [10, 20, 30])]),
                                      i32_table([("k2", [10, 10, 999]), ("b2", [1, 2, 3])]), JoinType.LeftSemi, k_filter())
    out = qref.execute(plan)
    assert out[0].schema.names == ["a1", "k1"]
    assert rows_of(out) == [(1, 10)]


def test_oracle_order_and_unmatched_batches():
    """build_join_indices (:237-271) walks the RIGHT rows in the outer loop; Full join appends unmatched left rows, then
    unmatched right rows, as a second batch (:168-226)."""
    left = i32_table([("a1", [1, 2, 3]), ("k1", [10, 20, 30])])
    right = i32_table([("k2", [20, 10, 10, 7]), ("b2", [1, 2, 3, 4])])
    inner = qref.execute(NestedLoopJoinExec.try_new(left, right, JoinType.Inner, k_filter()))
    assert rows_of(inner) == [(2, 20, 20, 1), (1, 10, 10, 2), (1, 10, 10, 3)]
    full = qref.execute(NestedLoopJoinExec.try_new(left, right, JoinType.Full, k_filter()))
    assert len(full) == 2
    assert rows_of(full) == [(2, 20, 20, 1), (1, 10, 10, 2), (1, 10, 10, 3), (3, 30, None, None), (None, None, 7, 4)]
    cross = qref.execute(NestedLoopJoinExec.try_new(left, right, JoinType.Inner, None))
    assert [r[0] for r in rows_of(cross)] == [1, 2, 3] * 4 and [r[3] for r in rows_of(cross)] == [1] * 3 + [2] * 3 + [3] * 3 + [4] * 3
    # empty right side (:86-119)
    empty = i32_table([("k2", []), ("b2", [])])
    assert qref.execute(NestedLoopJoinExec.try_new(left, empty, JoinType.Inner, None)) == []
    assert qref.execute(NestedLoopJoinExec.try_new(left, empty, JoinType.Right, None)) == []
    assert rows_of(qref.execute(NestedLoopJoinExec.try_new(left, empty, JoinType.Left, None))) == [(1, 10, None, None), (2, 20, None, None),
                                                                                                  (3, 30, None, None)]
    semi = qref.execute(NestedLoopJoinExec.try_new(left, empty, JoinType.LeftSemi, None))
    assert len(semi) == 1 and semi[0].num_rows == 0


# ---- GPU parity ---------------------------------------------------------------------------------------------
def _random_tables(rng, nl, nr, null_frac=0.15):
    def tab(n, names, splits):
        def col(lo, hi):
            vals = rng.integers(lo, hi, n)
            ok = rng.random(n) >= null_frac
            return pa.array([int(v) if o else None for v, o in zip(vals, ok)], pa.int32())
        cols = [col(0, 40), col(-100, 100)]
        strs = pa.array([None if rng.random() < null_frac else ["", "a", "BUILDING", "zz top"][int(v)] for v in rng.integers(0, 4, n)], pa.string())
        schema = pa.schema([(names[0], pa.int32()), (names[1], pa.int32()), (names[2], pa.string())])
        full = pa.record_batch(cols + [strs], schema=schema)
        batches, off = [], 0
        for s_ in splits:
            batches.append(full.slice(off, s_))
            off += s_
        return MemoryTable.try_new(schema, batches)
    return tab(nl, ("a", "b", "ls"), [nl // 3, 0, nl - nl // 3]), tab(nr, ("c", "d", "rs"), [nr])


def _same(name, plan, ctx, ordered=True):
    got, ref = plan.execute(ctx), qref.execute(plan)
    if got and ref:
        assert got[0].schema.types == ref[0].schema.types
        assert got[0].schema.names == ref[0].schema.names
    check_rows(name, rows_of(got), rows_of(ref), ordered=ordered)
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("join_type", JOIN_TYPES, ids=[j.name for j in JOIN_TYPES])
@pytest.mark.parametrize("with_filter", [False, True])
def test_gpu_nested_loop_join_random(gpu_ctx, join_type, with_filter):
    rng = np.random.default_rng(77)
    lt, rt = _random_tables(rng, 90, 70)
    jf = None
    if with_filter:  # non-equi condition over nullable columns: NULL drops the pair (join_filter_indices :292-297)
        schema = pa.schema([lt.schema.field("a"), rt.schema.field("c"), rt.schema.field("d")])
        expr = bx(bx(Column("a", 0), "Lt", Column("c", 1)), "And", bx(Column("d", 2), "Gt", Literal(ScalarValue.Int32(-50))))
        jf = JoinFilter(expr, schema, [(0, JoinSide.Left), (0, JoinSide.Right), (1, JoinSide.Right)])
    plan = NestedLoopJoinExec.try_new(Scan(lt.schema, lt, None, None), Scan(rt.schema, rt, None, None), join_type, jf)
    _same(f"nlj {join_type.name}", plan, gpu_ctx)
    assert "nested-loop-join" in plan.last_strategy()
    if join_type in (JoinType.LeftSemi, JoinType.LeftAnti):
        return
    # an aggregate on top (late materialisation of both sides, NULL-padded rows included)
    n_left = len(lt.schema)
    agg = HashAggregate(pa.schema([("ls", pa.string()), ("n", pa.int64()), ("sd", pa.int64())]), plan, [Column("ls", 2)],
                        [CountAggregateExpr(lit(1)), SumAggregateExpr(CastExpr(Column("d", n_left + 1), pa.int64()), pa.int64())])
    ref = qref.execute(agg)
    check_rows(f"nlj {join_type.name} -> aggregate", rows_of(agg.execute(gpu_ctx)), rows_of(ref), ordered=False)


@pytest.mark.gpu
@pytest.mark.parametrize("join_type", JOIN_TYPES, ids=[j.name for j in JOIN_TYPES])
def test_gpu_nested_loop_join_empty_sides(gpu_ctx, join_type):
    left = i32_table([("a1", [1, 2, 3]), ("k1", [10, 20, 30])])
    right = i32_table([("k2", [10, 10, 999]), ("b2", [1, 2, 3])])
    e_left = i32_table([("a1", []), ("k1", [])])
    e_right = i32_table([("k2", []), ("b2", [])])
    for name, l, r in (("empty right", left, e_right), ("empty left", e_left, right), ("both empty", e_left, e_right),
                       ("reference vectors", left, right)):
        for f in (None, k_filter()):
            plan = NestedLoopJoinExec.try_new(l, r, join_type, f)
            got = _same(f"nlj {join_type.name} {name}", plan, gpu_ctx)
            ref = qref.execute(plan)
            assert (len(got) == 0) == (len(ref) == 0), (name, join_type)   # Inner / Right over an empty right side: no batch at all


@pytest.mark.gpu
def test_gpu_nested_loop_join_chunked_cross_product(gpu_ctx):
    """the cross product is generated in chunks of right rows: results and order must not depend on the chunk size"""
    rng = np.random.default_rng(5)
    lt, rt = _random_tables(rng, 130, 61, null_frac=0.0)
    schema = pa.schema([lt.schema.field("b"), rt.schema.field("d")])
    jf = JoinFilter(bx(Column("b", 0), "GtEq", Column("d", 1)), schema, [(1, JoinSide.Left), (1, JoinSide.Right)])
    old = os.environ.get("QGPU_NLJ_MAX_PAIRS")
    try:
        for cap in ("1000", "131", "1"):
            os.environ["QGPU_NLJ_MAX_PAIRS"] = cap
            for jt in (JoinType.Inner, JoinType.Full, JoinType.LeftAnti):
                plan = NestedLoopJoinExec.try_new(Scan(lt.schema, lt, None, None), Scan(rt.schema, rt, None, None), jt, jf)
                _same(f"nlj chunk {cap} {jt.name}", plan, gpu_ctx)
    finally:
        if old is None:
            os.environ.pop("QGPU_NLJ_MAX_PAIRS", None)
        else:
            os.environ["QGPU_NLJ_MAX_PAIRS"] = old
