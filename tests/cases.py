"""Plan builders for the reference's own known-answer tests (tests/golden/reference_vectors.json).

Each case is (name, plan, expected_rows, ordered).  The same plan object is executed by the oracle
(oracle/qref.py, CPU) and by the GPU operators through the C ABI, and both are compared with the
reference's expected rows.  SQL cases are planned by hand the way qurious's planner does
(INT/BIGINT -> Int64, literals Int64/Float64, count(*) -> COUNT(Int64(1)), HAVING -> Filter above the
aggregate, pushed-down WHERE -> Scan filter).
"""
from __future__ import annotations

import datetime
import json
import os
from typing import Any, Dict, Iterator, List, Optional, Sequence, Tuple

import pyarrow as pa

from qurious_b200.datatypes import JoinSide, JoinType, Operator, ScalarValue
from qurious_b200.physical.expr import (AvgAggregateExpr, BinaryExpr, CastExpr, Column, CountAggregateExpr, IsNotNull,
                                        IsNull, Literal, MaxAggregateExpr, MinAggregateExpr, SumAggregateExpr)
from qurious_b200.physical.plan import (Filter, HashAggregate, HashJoinExec, JoinFilter, MemoryTable,
                                        NoGroupingAggregate, Projection, Scan)

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
I64 = pa.int64()
OPS = {o.name: o for o in Operator}


def lit(v) -> Literal:
    if isinstance(v, bool):
        return Literal(ScalarValue.Boolean(v))
    if isinstance(v, int):
        return Literal(ScalarValue.Int64(v))
    if isinstance(v, float):
        return Literal(ScalarValue.Float64(v))
    return Literal(ScalarValue.Utf8(v))


def bx(l, op: str, r) -> BinaryExpr:
    return BinaryExpr(l, OPS[op], r)


def table(cols: Dict[str, Sequence], types: Optional[Dict[str, pa.DataType]] = None, default=I64,
          splits: Optional[Sequence[int]] = None, nullable: bool = True) -> MemoryTable:
    """An in-memory table; `splits` = row counts of the individual INSERT batches."""
    types = types or {}
    fields = [pa.field(k, types.get(k, default), nullable) for k in cols]
    schema = pa.schema(fields)
    n = len(next(iter(cols.values()))) if cols else 0
    arrays = [pa.array(v, type=f.type) for v, f in zip(cols.values(), fields)]
    full = pa.record_batch(arrays, schema=schema)
    if splits is None:
        splits = [n] if n > 0 else []
    batches, off = [], 0
    for s in splits:
        batches.append(full.slice(off, s))
        off += s
    return MemoryTable.try_new(schema, batches)


def scan(t: MemoryTable, filter=None) -> Scan:
    return Scan(t.schema, t, None, filter)


def col(t, name: str) -> Column:
    schema = t.schema
    return Column(name, schema.get_field_index(name))


def schema_of(*fields: Tuple[str, pa.DataType]) -> pa.Schema:
    return pa.schema([pa.field(n, t, True) for n, t in fields])


def rows_of(batches: Sequence[pa.RecordBatch]) -> List[tuple]:
    out: List[tuple] = []
    for b in batches:
        cols = [c.to_pylist() for c in b.columns]
        for i in range(b.num_rows):
            out.append(tuple(c[i] for c in cols))
    return out


def norm(v: Any):
    """Normalise a cell for comparison: dates -> iso strings, Decimal -> (unscaled int)."""
    if isinstance(v, (datetime.date, datetime.datetime)):
        return v.isoformat()
    return v


def sort_key(row: tuple):
    return tuple((x is None, x if x is not None else 0) for x in row)


Case = Tuple[str, Any, List[tuple], bool]


def _scalar_rows(v) -> List[tuple]:
    return [(v,)]


def golden_cases() -> Iterator[Case]:
    G = GOLDEN
    # ---------------------------------------------------------------- basic_test.slt: EXTRACT on a date literal
    from qurious_b200.physical.expr import DatetimeExtract, Function
    ex = G["extract"]
    one = table({"x": [0]}, nullable=False)
    d_lit = CastExpr(Literal(ScalarValue.Utf8(ex["date"])), pa.date32())
    yield ("basic_test.slt:extract",
           Projection(schema_of(("year", I64), ("month", I64), ("day", I64)), scan(one),
                      [Function(DatetimeExtract(), [Literal(ScalarValue.Utf8(p_)), d_lit]) for p_ in ("YEAR", "MONTH", "DAY")]),
           [(ex["year"], ex["month"], ex["day"])], True)
    # ---------------------------------------------------------------- sort.rs / limit.rs unit tests, order_by.slt, limit.slt
    from qurious_b200.physical.plan import Limit, PhyscialSortExpr, Sort, SortOptions
    SL = G["sort_limit"]
    PT = {"int32": pa.int32(), "float64": pa.float64(), "uint64": pa.uint64()}

    def sort_keys(t, keys):
        return [PhyscialSortExpr(col(t, k), SortOptions(descending=desc, nulls_first=nf)) for k, desc, nf in keys]
    for c in SL["sort_unit"]:
        t = table(c["table"], types={k: PT[v] for k, v in c["types"].items()}, nullable=False)
        yield (f"sort:{c['name']}", Sort(sort_keys(t, c["keys"]), scan(t), c["limit"]), [tuple(r) for r in c["expected"]], True)
    lu = SL["limit_unit"]
    t = table(lu["table"], types={k: PT[v] for k, v in lu["types"].items()}, nullable=False)
    yield ("limit:test_limit", Limit(scan(t), lu["fetch"], lu["skip"]), [tuple(r) for r in lu["expected"]], True)
    for n_, c in enumerate(SL["order_by_slt"]):
        t = table(c["table"])
        sel = Projection(schema_of(*[(k, I64) for k in c["select"]]), scan(t), [col(t, k) for k in c["select"]])
        plan = Sort([PhyscialSortExpr(Column(k, c["select"].index(k)), SortOptions(descending=desc, nulls_first=nf)) for k, desc, nf in c["keys"]],
                    sel)
        yield (f"order_by.slt:{n_}", plan, [tuple(r) for r in c["expected"]], True)
    ls = SL["limit_slt"]
    for n_, c in enumerate(ls["cases"]):
        t = table(ls["table"], nullable=False)
        plan = Limit(Projection(schema_of(("v1", I64)), scan(t), [col(t, "v1")]), c["fetch"], c["skip"])
        yield (f"limit.slt:{n_}", plan, [(x,) for x in c["expected"]], True)
    # ---------------------------------------------------------------- binary.rs unit tests
    for kind, dt, out_t in (("binary_comparison", pa.int32(), pa.bool_()), ("binary_arithmetic", pa.int32(), pa.int32()),
                            ("binary_logical", pa.bool_(), pa.bool_())):
        for c in G[kind]["cases"]:
            t = table({"left": c["left"], "right": c["right"]}, default=dt, nullable=False)
            e = bx(Column("left", 0), c["op"], Column("right", 1))
            plan = Projection(schema_of(("r", out_t)), scan(t), [e])
            yield (f"{kind}:{c['op']}", plan, [(x,) for x in c["expected"]], True)
    d = G["decimal_nested"]
    dt = pa.decimal128(*d["column_type"])
    import decimal as _dec

    def dec(raw, scale):
        return _dec.Decimal(raw).scaleb(-scale)
    t = table({"l_extendedprice": [dec(x, 2) for x in d["l_extendedprice_raw"]],
               "l_discount": [dec(x, 2) for x in d["l_discount_raw"]]}, default=dt)
    one_minus = bx(CastExpr(Literal(ScalarValue.Int16(1)), dt), "Sub", Column("l_discount", 1))
    e = bx(Column("l_extendedprice", 0), "Mul", one_minus)
    et = pa.decimal128(*d["expected_type"])
    plan = Projection(schema_of(("r", et)), scan(t), [e])
    yield ("decimal_nested", plan, [(dec(x, d["expected_type"][1]),) for x in d["expected_raw"]], True)

    # ---------------------------------------------------------------- hash_join.rs unit tests
    for name, c in G["hash_join"]["cases"].items():
        lt = table(c["left"], default=pa.int32()) if any(len(v) for v in c["left"].values()) else \
            table_with_empty_batch(c["left"], pa.int32())
        rt = table(c["right"], default=pa.int32()) if any(len(v) for v in c["right"].values()) else \
            table_with_empty_batch(c["right"], pa.int32())
        on = [(Column(list(c["left"])[li], li), Column(list(c["right"])[ri], ri)) for li, ri in c["on"]]
        plan = HashJoinExec.try_new(scan(lt), scan(rt), JoinType[c["join_type"]], on, None)
        yield (f"hash_join:{name}", plan, [tuple(r) for r in c["expected"]], True)

    # ---------------------------------------------------------------- nest_loop_join.rs unit tests
    from qurious_b200.datatypes import JoinSide
    from qurious_b200.physical.plan import JoinFilter, NestedLoopJoinExec
    for name, c in G["nested_loop_join"]["cases"].items():
        lt = table(c["left"], default=pa.int32()) if any(len(v) for v in c["left"].values()) else \
            table_with_empty_batch(c["left"], pa.int32())
        rt = table(c["right"], default=pa.int32()) if any(len(v) for v in c["right"].values()) else \
            table_with_empty_batch(c["right"], pa.int32())
        jf = None
        if c["filter"]:
            fs = pa.schema([pa.field("k1", pa.int32(), False), pa.field("k2", pa.int32(), False)])
            jf = JoinFilter(bx(Column("k1", 0), "Eq", Column("k2", 1)), fs, [(1, JoinSide.Left), (0, JoinSide.Right)])
        plan = NestedLoopJoinExec.try_new(scan(lt), scan(rt), JoinType[c["join_type"]], jf)
        yield (f"nested_loop_join:{name}", plan, [tuple(r) for r in c["expected"]], True)
    # ---------------------------------------------------------------- cross_join.rs unit test
    from qurious_b200.physical.plan import CrossJoin
    cj = G["cross_join"]
    yield ("cross_join:test_cross_join", CrossJoin.new(scan(table(cj["left"], default=pa.int32())), scan(table(cj["right"], default=pa.int32()))),
           [tuple(r) for r in cj["expected"]], True)
    # ---------------------------------------------------------------- aggregate/hash.rs unit test
    hu = G["hash_aggregate_unit"]
    t = table(hu["table"], default=pa.int32(), nullable=False)
    yield ("hash_aggregate:test_group_by",
           HashAggregate(schema_of(("c1", pa.int32()), ("b1", pa.int32()), ("MAX(a1)", pa.int32())), scan(t),
                         [col(t, k) for k in hu["group_by"]], [MaxAggregateExpr(col(t, "a1"), pa.int32())]),
           [tuple(r) for r in hu["expected"]], False)

    S = G["slt"]
    # ---------------------------------------------------------------- aggregation.slt
    a = S["aggregation_t1"]
    for tag, splits in (("one_insert", None), ("four_inserts", [1, 1, 1, 1])):
        t = table(a["table"], types={"v3": pa.float64()}, splits=splits, nullable=False)
        v1, v2, v3 = col(t, "v1"), col(t, "v2"), col(t, "v3")
        p = NoGroupingAggregate(schema_of(("SUM(v1)", I64), ("SUM(v2)", I64)), scan(t),
                                [SumAggregateExpr(v1, I64), SumAggregateExpr(v2, I64)])
        yield (f"aggregation[{tag}]:sum(v1)+sum(v2)",
               Projection(schema_of(("r", I64)), p, [bx(Column("a", 0), "Add", Column("b", 1))]),
               _scalar_rows(a["sum_v1_plus_sum_v2"]), True)
        yield (f"aggregation[{tag}]:sum(v1),sum(v3)",
               NoGroupingAggregate(schema_of(("SUM(v1)", I64), ("SUM(v3)", pa.float64())), scan(t),
                                   [SumAggregateExpr(v1, I64), SumAggregateExpr(v3, pa.float64())]),
               [(a["sum_v1"], a["sum_v3"])], True)
        yield (f"aggregation[{tag}]:min(v1)",
               NoGroupingAggregate(schema_of(("MIN(v1)", I64)), scan(t), [MinAggregateExpr(v1, I64)]),
               _scalar_rows(a["min_v1"]), True)
        yield (f"aggregation[{tag}]:max(v1)",
               NoGroupingAggregate(schema_of(("MAX(v1)", I64)), scan(t), [MaxAggregateExpr(v1, I64)]),
               _scalar_rows(a["max_v1"]), True)
        yield (f"aggregation[{tag}]:max(v1) where v2>3",
               NoGroupingAggregate(schema_of(("MAX(v1)", I64)), scan(t, bx(v2, "Gt", lit(3))), [MaxAggregateExpr(v1, I64)]),
               _scalar_rows(a["max_v1_where_v2_gt_3"]), True)
        yield (f"aggregation[{tag}]:count(v1)",
               NoGroupingAggregate(schema_of(("COUNT(v1)", I64)), scan(t), [CountAggregateExpr(v1)]),
               _scalar_rows(a["count_v1"]), True)
        agg = HashAggregate(schema_of(("v2", I64), ("SUM(v1)", I64)), scan(t), [v2], [SumAggregateExpr(v1, I64)])
        yield (f"aggregation[{tag}]:sum(v1),v2 group by v2",
               Projection(schema_of(("SUM(v1)", I64), ("v2", I64)), agg, [Column("SUM(v1)", 1), Column("v2", 0)]),
               [tuple(r) for r in a["sum_v1_group_by_v2"]], False)
    e = S["aggregation_empty"]
    empty = table({"v1": []}, nullable=False)  # CREATE TABLE without INSERT: zero batches
    yield ("aggregation_empty:count(v1)",
           NoGroupingAggregate(schema_of(("COUNT(v1)", I64)), scan(empty), [CountAggregateExpr(col(empty, "v1"))]),
           _scalar_rows(e["count_v1_empty"]), True)
    yield ("aggregation_empty:count(v1) group by v1",
           HashAggregate(schema_of(("v1", I64), ("COUNT(v1)", I64)), scan(empty), [col(empty, "v1")],
                         [CountAggregateExpr(col(empty, "v1"))]), [], True)
    yield ("aggregation_empty:sum(x)",
           NoGroupingAggregate(schema_of(("SUM(x)", I64)), scan(empty), [SumAggregateExpr(col(empty, "v1"), I64)]),
           _scalar_rows(e["sum_x_empty"]), True)

    # ---------------------------------------------------------------- group_by.slt
    g = S["group_by"]
    t = table(g["table"])
    v1, v2 = col(t, "v1"), col(t, "v2")
    v2p1 = bx(v2, "Add", lit(1))
    yield ("group_by:v2+1,sum(v1)",
           HashAggregate(schema_of(("v2 + 1", I64), ("SUM(v1)", I64)), scan(t), [v2p1], [SumAggregateExpr(v1, I64)]),
           [tuple(r) for r in g["v2p1_sum_v1"]], False)
    agg = HashAggregate(schema_of(("a", I64), ("SUM(v1)", I64), ("COUNT(1)", I64)), scan(t), [v2p1],
                        [SumAggregateExpr(v1, I64), CountAggregateExpr(lit(1))])
    yield ("group_by:sum(v1),a,count(*)",
           Projection(schema_of(("SUM(v1)", I64), ("a", I64), ("COUNT(1)", I64)), agg,
                      [Column("SUM(v1)", 1), Column("a", 0), Column("COUNT(1)", 2)]),
           [tuple(r) for r in g["sum_v1_v2p1_count"]], False)
    agg = HashAggregate(schema_of(("v2 + 1", I64), ("v2", I64), ("SUM(v1)", I64)), scan(t), [v2p1, v2],
                        [SumAggregateExpr(v1, I64)])
    yield ("group_by:v2,v2+1,sum(v1) group by v2+1,v2",
           Projection(schema_of(("v2", I64), ("v2 + 1", I64), ("SUM(v1)", I64)), agg,
                      [Column("v2", 1), Column("v2 + 1", 0), Column("SUM(v1)", 2)]),
           [tuple(r) for r in g["v2_v2p1_sum_v1"]], False)
    agg = HashAggregate(schema_of(("v1 + 1", I64), ("COUNT(1)", I64)), scan(t), [bx(v1, "Add", lit(1))],
                        [CountAggregateExpr(lit(1))])
    yield ("group_by:v1+1+count(*)",
           Projection(schema_of(("r", I64)), agg, [bx(Column("v1 + 1", 0), "Add", Column("COUNT(1)", 1))]),
           [(x,) for x in g["v1p1_plus_count"]], False)

    # ---------------------------------------------------------------- having.slt
    h = S["having"]
    t = table(h["table"])
    x, y = col(t, "x"), col(t, "y")
    agg = HashAggregate(schema_of(("b", I64), ("SUM(x)", I64)), scan(t), [y], [SumAggregateExpr(x, I64)])
    yield ("having:b,sum having b=2", Filter(agg, bx(Column("b", 0), "Eq", lit(2))),
           [tuple(r) for r in h["b_sum_having_b_eq_2"]], True)
    agg = HashAggregate(schema_of(("b", I64), ("a", I64)), scan(t), [y], [CountAggregateExpr(x)])
    yield ("having:count(x) a,y b having a>1",
           Projection(schema_of(("a", I64), ("b", I64)), Filter(agg, bx(Column("a", 1), "Gt", lit(1))),
                      [Column("a", 1), Column("b", 0)]),
           [tuple(r) for r in h["count_x_y_having_a_gt_1"]], True)
    agg = HashAggregate(schema_of(("b", I64), ("a", I64)), scan(t), [bx(y, "Add", lit(1))], [CountAggregateExpr(x)])
    yield ("having:count(x),y+1 having b+1=24",
           Projection(schema_of(("a", I64), ("b", I64)),
                      Filter(agg, bx(bx(Column("b", 0), "Add", lit(1)), "Eq", lit(24))), [Column("a", 1), Column("b", 0)]),
           [tuple(r) for r in h["count_x_yp1_having_bp1_eq_24"]], True)
    agg = HashAggregate(schema_of(("x", I64), ("MAX(y)", I64)), scan(t), [x], [MaxAggregateExpr(y, I64)])
    yield ("having:x having max(y)=22",
           Projection(schema_of(("x", I64)), Filter(agg, bx(Column("MAX(y)", 1), "Eq", lit(22))), [Column("x", 0)]),
           [tuple(r) for r in h["x_having_max_y_eq_22"]], True)

    # ---------------------------------------------------------------- count.slt
    c = S["count"]
    t = table(c["table"])
    v = col(t, "v")
    cnt = lambda src: NoGroupingAggregate(schema_of(("COUNT(1)", I64)), src, [CountAggregateExpr(lit(1))])  # noqa: E731
    yield ("count:count(*)", cnt(scan(t)), _scalar_rows(c["count_star"]), True)
    yield ("count:count(*) where v>5", cnt(scan(t, bx(v, "Gt", lit(5)))), _scalar_rows(c["count_where_v_gt_5"]), True)
    p = NoGroupingAggregate(schema_of(("COUNT(1)", I64), ("MIN(v)", I64)), scan(t),
                            [CountAggregateExpr(lit(1)), MinAggregateExpr(v, I64)])
    yield ("count:count(*)+min(v)", Projection(schema_of(("r", I64)), p, [bx(Column("a", 0), "Add", Column("b", 1))]),
           _scalar_rows(c["count_plus_min"]), True)
    t7 = table({"v": [x_ for x_ in c["table"]["v"] if x_ != 7]})
    yield ("count:after delete v=7", cnt(scan(t7, bx(col(t7, "v"), "Gt", lit(5)))),
           _scalar_rows(c["after_delete_v_eq_7_count_where_v_gt_5"]), True)
    yield ("count:where 0=1", cnt(scan(t, bx(lit(0), "Eq", lit(1)))), _scalar_rows(c["count_where_false"]), True)

    # ---------------------------------------------------------------- bigint.slt
    b = S["bigint"]
    t = table(b["table"])
    v2 = col(t, "v2")
    gt2 = bx(v2, "Gt", lit(2))
    yield ("bigint:count(v2) where v2>2",
           NoGroupingAggregate(schema_of(("c", I64)), scan(t, gt2), [CountAggregateExpr(v2)]), _scalar_rows(b["count_v2_gt_2"]), True)
    yield ("bigint:min(v2) where v2>2",
           NoGroupingAggregate(schema_of(("c", I64)), scan(t, gt2), [MinAggregateExpr(v2, I64)]), _scalar_rows(b["min_v2_gt_2"]), True)
    yield ("bigint:max(v2) where v2>2",
           NoGroupingAggregate(schema_of(("c", I64)), scan(t, gt2), [MaxAggregateExpr(v2, I64)]), _scalar_rows(b["max_v2_gt_2"]), True)
    yield ("bigint:sum(v2) where v2<10",
           NoGroupingAggregate(schema_of(("c", I64)), scan(t, bx(v2, "Lt", lit(10))), [SumAggregateExpr(v2, I64)]),
           _scalar_rows(b["sum_v2_lt_10"]), True)

    # ---------------------------------------------------------------- filter.slt
    f = S["filter"]
    t = table(f["table1"], nullable=False)
    v1, v2 = col(t, "v1"), col(t, "v2")
    one = lambda src, name: Projection(schema_of((name, I64)), src, [Column(name, {"v1": 0, "v2": 1}[name])])  # noqa: E731
    yield ("filter:v1 where v1>2", one(scan(t, bx(v1, "Gt", lit(2))), "v1"), [(x_,) for x_ in f["t1_v1_where_v1_gt_2"]], False)
    yield ("filter:v2 where 3>v1", one(scan(t, bx(lit(3), "Gt", v1)), "v2"), [(x_,) for x_ in f["t1_v2_where_3_gt_v1"]], True)
    t = table(f["table2"], nullable=False, splits=[7, 7])
    v1, v2 = col(t, "v1"), col(t, "v2")
    AND, OR = (lambda l, r: bx(l, "And", r)), (lambda l, r: bx(l, "Or", r))
    q = {
        "t2_q1": AND(bx(v1, "Gt", lit(2)), bx(v1, "Lt", lit(4))),
        "t2_q2": AND(OR(bx(lit(-7), "Lt", v1), bx(lit(9), "LtEq", v1)), bx(v1, "Eq", lit(3))),
        "t2_q3": OR(AND(bx(lit(-8), "Lt", v1), bx(v1, "LtEq", lit(-7))), AND(bx(v1, "GtEq", lit(1)), bx(lit(2), "Gt", v1))),
        "t2_q4": AND(OR(AND(bx(v1, "GtEq", lit(-8)), bx(lit(-4), "GtEq", v1)), AND(bx(v1, "GtEq", lit(0)), bx(lit(5), "Gt", v1))),
                     OR(AND(bx(v1, "Gt", lit(0)), bx(v1, "LtEq", lit(1))), AND(bx(v1, "Gt", lit(-8)), bx(v1, "Lt", lit(-6))))),
        "t2_q5": AND(OR(bx(lit(-7), "Lt", v1), bx(lit(9), "LtEq", v1)), bx(v2, "Eq", lit(3))),
        "t2_q6": OR(AND(bx(lit(-8), "Lt", v1), bx(v2, "LtEq", lit(-7))), AND(bx(v1, "GtEq", lit(1)), bx(lit(2), "Gt", v2))),
        "t2_q7": AND(OR(AND(bx(v2, "GtEq", lit(-8)), bx(lit(-4), "GtEq", v1)), AND(bx(v1, "GtEq", lit(0)), bx(lit(5), "Gt", v2))),
                     OR(AND(bx(v2, "Gt", lit(0)), bx(v1, "LtEq", lit(1))), AND(bx(v1, "Gt", lit(-8)), bx(v2, "Lt", lit(-6))))),
    }
    for name, pred in q.items():
        out_col = "v1" if name == "t2_q1" else "v2"
        yield (f"filter:{name}", one(scan(t, pred), out_col), [(x_,) for x_ in f[name]], False)

    # ---------------------------------------------------------------- filter_null.slt
    fn = S["filter_null"]
    t = table(fn["table1"])
    yield ("filter_null:t1 v1>1", scan(t, bx(col(t, "v1"), "Gt", lit(1))), [tuple(r) for r in fn["t1_where_v1_gt_1"]], True)
    yield ("filter_null:t1 v1<2", scan(t, bx(col(t, "v1"), "Lt", lit(2))), [tuple(r) for r in fn["t1_where_v1_lt_2"]], True)
    t = table(fn["table2"])
    yield ("filter_null:t2 v1>1 (NULL row dropped)", scan(t, bx(col(t, "v1"), "Gt", lit(1))),
           [tuple(r) for r in fn["t2_where_v1_gt_1"]], True)
    yield ("filter_null:t2 Filter operator", Filter(scan(t), bx(col(t, "v1"), "Gt", lit(1))),
           [tuple(r) for r in fn["t2_where_v1_gt_1"]], True)

    # ---------------------------------------------------------------- where.slt
    w = S["where"]
    t = table(w["table1"], nullable=False)
    v1, v2 = col(t, "v1"), col(t, "v2")
    yield ("where:v1>v2", scan(t, bx(v1, "Gt", v2)), [tuple(r) for r in w["t1_v1_gt_v2"]], True)
    yield ("where:v2>2", scan(t, bx(v2, "Gt", lit(2))), [tuple(r) for r in w["t1_v2_gt_2"]], True)
    yield ("where:v1=1 or v2=2", scan(t, bx(bx(v1, "Eq", lit(1)), "Or", bx(v2, "Eq", lit(2)))),
           [tuple(r) for r in w["t1_v1_eq_1_or_v2_eq_2"]], True)
    yield ("where:v1=1 and v2=1", scan(t, bx(bx(v1, "Eq", lit(1)), "And", bx(v2, "Eq", lit(1)))),
           [tuple(r) for r in w["t1_v1_eq_1_and_v2_eq_1"]], True)
    sumv2 = lambda src: NoGroupingAggregate(schema_of(("SUM(v2)", I64)), src, [SumAggregateExpr(Column("v2", 1), I64)])  # noqa: E731
    yield ("where:sum(v2) v1!=1", sumv2(scan(t, bx(v1, "NotEq", lit(1)))), _scalar_rows(w["t1_sum_v2_where_v1_ne_1"]), True)
    t = table(w["table2"], nullable=False)
    v1 = col(t, "v1")
    yield ("where:sum(v2) v1<1", sumv2(scan(t, bx(v1, "Lt", lit(1)))), _scalar_rows(w["t2_sum_v2_v1_lt_1"]), True)
    yield ("where:sum(v2) v1<=1", sumv2(scan(t, bx(v1, "LtEq", lit(1)))), _scalar_rows(w["t2_sum_v2_v1_le_1"]), True)
    yield ("where:sum(v2) v1>=1", sumv2(scan(t, bx(v1, "GtEq", lit(1)))), _scalar_rows(w["t2_sum_v2_v1_ge_1"]), True)
    t = table(w["table3"])
    v1 = col(t, "v1")
    v2only = lambda src: Projection(schema_of(("v2", I64)), src, [Column("v2", 1)])  # noqa: E731
    yield ("where:v1 is null", v2only(scan(t, IsNull(v1))), [(x_,) for x_ in w["t3_v2_where_v1_is_null"]], True)
    yield ("where:v1 is not null", v2only(scan(t, IsNotNull(v1))), [(x_,) for x_ in w["t3_v2_where_v1_is_not_null"]], True)

    # ---------------------------------------------------------------- join.slt
    j = S["join"]
    x_t, y0, y_t = table(j["x"]), table(j["y_empty"]), table(j["y"])
    on_ac = lambda l, r: [(col(l, "a"), col(r, "c"))]  # noqa: E731
    yield ("join:x join y(empty)", HashJoinExec.try_new(scan(x_t), scan(y0), JoinType.Inner, on_ac(x_t, y0), None),
           [tuple(r) for r in j["x_join_y_empty"]], True)
    yield ("join:x join y", HashJoinExec.try_new(scan(x_t), scan(y_t), JoinType.Inner, on_ac(x_t, y_t), None),
           [tuple(r) for r in j["x_join_y"]], True)
    a_t, b0, b_t = table(j["a"]), table(j["b_empty"]), table(j["b"])
    on13 = lambda l, r: [(col(l, "v1"), col(r, "v3"))]  # noqa: E731
    yield ("join:a left join b(empty)", HashJoinExec.try_new(scan(a_t), scan(b0), JoinType.Left, on13(a_t, b0), None),
           [tuple(r) for r in j["a_left_b_empty"]], True)
    yield ("join:a left join b", HashJoinExec.try_new(scan(a_t), scan(b_t), JoinType.Left, on13(a_t, b_t), None),
           [tuple(r) for r in j["a_left_b"]], True)
    yield ("join:a right join b", HashJoinExec.try_new(scan(a_t), scan(b_t), JoinType.Right, on13(a_t, b_t), None),
           [tuple(r) for r in j["a_right_b"]], True)
    yield ("join:a full join b", HashJoinExec.try_new(scan(a_t), scan(b_t), JoinType.Full, on13(a_t, b_t), None),
           [tuple(r) for r in j["a_full_b"]], True)
    b3 = table(j["b3"])
    on2 = [(col(a_t, "v1"), col(b3, "v3")), (col(a_t, "v2"), col(b3, "v4"))]
    yield ("join:a join b3 on two keys", HashJoinExec.try_new(scan(a_t), scan(b3), JoinType.Inner, on2, None),
           [tuple(r) for r in j["a_join_b3_two_keys"]], True)
    jf = JoinFilter(bx(Column("v1", 0), "Lt", Column("v5", 1)), schema_of(("v1", I64), ("v5", I64)),
                    [(0, JoinSide.Left), (2, JoinSide.Right)])
    yield ("join:a join b3 on two keys and v1<v5", HashJoinExec.try_new(scan(a_t), scan(b3), JoinType.Inner, on2, jf),
           [tuple(r) for r in j["a_join_b3_two_keys_filter_v1_lt_v5"]], True)

    # ---------------------------------------------------------------- type.slt
    td = S["type_date"]
    t = table({"v1": [datetime.date.fromisoformat(s) for s in td["table"]["v1"]]}, default=pa.date32(), nullable=False)
    pred = bx(col(t, "v1"), "Lt", CastExpr(lit("2021-01-01"), pa.date32()))
    yield ("type:date < date literal", scan(t, pred), [(s,) for s in td["expected"]], True)
    ts = S["type_smallint"]
    t = table(ts["table"], default=pa.int16(), nullable=False)
    a_ = col(t, "a")
    plan = Projection(schema_of(("p", pa.int16()), ("m", pa.int16()), ("x", pa.int16()), ("d", pa.int16())), scan(t),
                      [bx(a_, "Add", a_), bx(a_, "Sub", a_), bx(a_, "Mul", a_), bx(a_, "Div", a_)])
    yield ("type:smallint arithmetic", plan, [tuple(r) for r in ts["expected"]], True)


def table_with_empty_batch(cols: Dict[str, Sequence], dt: pa.DataType) -> MemoryTable:
    """build_table_scan_i32 with empty vectors: ONE RecordBatch of zero rows (test_utils.rs:218-236)."""
    schema = pa.schema([pa.field(k, dt, True) for k in cols])
    batch = pa.record_batch([pa.array([], type=dt) for _ in cols], schema=schema)
    return MemoryTable.try_new(schema, [batch])


def check_rows(name: str, got: List[tuple], expected: List[tuple], ordered: bool):
    got = [tuple(norm(v) for v in r) for r in got]
    expected = [tuple(norm(v) for v in r) for r in expected]
    if not ordered:
        got, expected = sorted(got, key=sort_key), sorted(expected, key=sort_key)
    assert len(got) == len(expected), f"{name}: row count {len(got)} != {len(expected)}\n got={got}\n exp={expected}"
    for g, e in zip(got, expected):
        assert len(g) == len(e), f"{name}: arity {g} vs {e}"
        for a, b in zip(g, e):
            if isinstance(b, float) and a is not None and (b != b or b in (float("inf"), float("-inf"))):
                import math
                assert isinstance(a, float) and ((math.isnan(a) and math.isnan(b)) or a == b), f"{name}: {g} != {e}"
            elif isinstance(b, float) and a is not None:
                assert abs(a - b) <= 1e-12 * max(1.0, abs(b)), f"{name}: {g} != {e}"
            else:
                assert a == b, f"{name}: {g} != {e}\n got={got}\n exp={expected}"
