"""TPC-H Q2, Q4, Q5, Q7-Q12 as the physical plans qurious builds for them ("widen to whole queries"), over
the eight synthetic tables.  Together with tpch.py's Q1 / Q3 / Q6 these are the 12 statements of qurious/tests/tpch/*.slt.

Plan shapes follow the reference's pipeline (planner/sql.rs -> the 8 optimizer rules -> planner/mod.rs:40-153):
  * `FROM a, b, c ...` is a left-deep chain in FROM order; a link becomes an Inner HashJoinExec when the WHERE clause holds
    an equi-condition between the accumulated left side and the next table, otherwise it STAYS a CrossJoin
    (optimizer/rule/eliminate_cross_join.rs:53-79) -- Q2, Q8 and Q9 begin with `part, supplier`, which nothing links;
  * single-table predicates are pushed into the Scan (pushdown_filter.rs; MemoryTable::scan evaluates them, memory.rs:79-95),
    predicates over several tables stay in a Filter above the join chain (Q7's nation pair);
  * operand types are unified by CastExprs exactly as utils/type_coercion.rs:37-170 + optimizer/rule/type_coercion.rs
    prescribe (Int64 next to a decimal -> Decimal128(20,0); decimal `/` -> both sides Float64; Date32 vs Utf8 literal ->
    CAST(Utf8 AS Date32); CASE branches -> their common type);
  * a correlated scalar subquery becomes a LEFT join against the grouped subquery followed by a Filter
    (scalar_subquery_to_join.rs:38-96: Q2), an uncorrelated one a LEFT NestedLoopJoin with the filter `true`
    (:86-88 + planner/mod.rs:316-320: Q11); EXISTS becomes a LeftSemi join (decorrelate_predicate_subquery.rs: Q4);
  * ORDER BY ... LIMIT n is Limit(Sort.new_with_limit(n)) (planner/mod.rs:69-83), a SubqueryAlias a column Projection.
The data is synthetic (tpch.py's counter-based generator, dbgen's vocabularies), so the VALUES differ from the goldens
in q*.slt; tests compare the GPU result of each plan with the oracle's (tests/test_gpu_tpch_queries.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import pyarrow as pa
import torch

from . import tpch
from .datatypes import JoinSide, JoinType, Operator, ScalarValue
from .physical.expr import (BinaryExpr, CaseExpr, CastExpr, Column, CountAggregateExpr, DatetimeExtract, Function, Like, Literal,
                            MinAggregateExpr, PhysicalExpr, SumAggregateExpr)
from .physical.plan import (CrossJoin, Filter, HashAggregate, HashJoinExec, JoinFilter, Limit, MemoryTable, NestedLoopJoinExec,
                            PhyscialSortExpr, Projection, Scan, Sort, SortOptions)
from .tpch import DEC, RawTable, rnd, uniform

O = Operator

# ------------------------------------------------------------------------------------------------
# the five remaining tables (tests/tpch/create_tables.slt:1-40,85-101; BIGINT / INTEGER -> Int64)
# ------------------------------------------------------------------------------------------------
NATIONS = [("ALGERIA", 0), ("ARGENTINA", 1), ("BRAZIL", 1), ("CANADA", 1), ("EGYPT", 4), ("ETHIOPIA", 0), ("FRANCE", 3),
           ("GERMANY", 3), ("INDIA", 2), ("INDONESIA", 2), ("IRAN", 4), ("IRAQ", 4), ("JAPAN", 2), ("JORDAN", 4), ("KENYA", 0),
           ("MOROCCO", 0), ("MOZAMBIQUE", 0), ("PERU", 1), ("CHINA", 2), ("ROMANIA", 3), ("SAUDI ARABIA", 4), ("VIETNAM", 2),
           ("RUSSIA", 3), ("UNITED KINGDOM", 3), ("UNITED STATES", 1)]
REGIONS = ["AFRICA", "AMERICA", "ASIA", "EUROPE", "MIDDLE EAST"]
COLORS = ["almond", "antique", "aquamarine", "azure", "beige", "bisque", "black", "blanched", "blue", "blush", "brown",
          "burlywood", "burnished", "chartreuse", "chiffon", "chocolate", "coral", "cornflower", "cornsilk", "cream", "cyan",
          "dark", "deep", "dim", "dodger", "drab", "firebrick", "floral", "forest", "frosted", "gainsboro", "ghost", "goldenrod",
          "green", "grey", "honeydew", "hot", "indian", "ivory", "khaki"]
TYPE_S1 = ["STANDARD", "SMALL", "MEDIUM", "LARGE", "ECONOMY", "PROMO"]
TYPE_S2 = ["ANODIZED", "BURNISHED", "PLATED", "POLISHED", "BRUSHED"]
TYPE_S3 = ["TIN", "NICKEL", "BRASS", "STEEL", "COPPER"]
P_TYPES = [f"{a} {b} {c}" for a in TYPE_S1 for b in TYPE_S2 for c in TYPE_S3]
CONTAINERS = [f"{a} {b}" for a in ("SM", "LG", "MED", "JUMBO", "WRAP") for b in ("CASE", "BOX", "BAG", "JAR", "PKG", "PACK", "CAN", "DRUM")]
P_NAMES = [" ".join(COLORS[(i * 7 + j * 11 + (i // 40) * 3) % len(COLORS)] for j in range(5)) for i in range(400)]

SUPPLIER_SCHEMA = pa.schema([("s_suppkey", pa.int64()), ("s_name", pa.string()), ("s_address", pa.string()), ("s_nationkey", pa.int64()),
                             ("s_phone", pa.string()), ("s_acctbal", DEC), ("s_comment", pa.string()), ("s_rev", pa.string())])
PART_SCHEMA = pa.schema([("p_partkey", pa.int64()), ("p_name", pa.string()), ("p_mfgr", pa.string()), ("p_brand", pa.string()),
                         ("p_type", pa.string()), ("p_size", pa.int64()), ("p_container", pa.string()), ("p_retailprice", DEC),
                         ("p_comment", pa.string()), ("p_rev", pa.string())])
PARTSUPP_SCHEMA = pa.schema([("ps_partkey", pa.int64()), ("ps_suppkey", pa.int64()), ("ps_availqty", pa.int64()),
                             ("ps_supplycost", DEC), ("ps_comment", pa.string()), ("ps_rev", pa.string())])
NATION_SCHEMA = pa.schema([("n_nationkey", pa.int64()), ("n_name", pa.string()), ("n_regionkey", pa.int64()), ("n_comment", pa.string()),
                           ("n_rev", pa.string())])
REGION_SCHEMA = pa.schema([("r_regionkey", pa.int64()), ("r_name", pa.string()), ("r_comment", pa.string()), ("r_rev", pa.string())])


def gen_supplier(sf: float, device="cpu") -> RawTable:
    n = tpch.n_suppliers(sf)
    i = torch.arange(n, dtype=torch.int64, device=device)
    cols = {"s_suppkey": i + 1, "s_nationkey": uniform(60, i, 0, 24), "s_acctbal": uniform(61, i, -99999, 999999)}
    names = [f"Supplier#{j + 1:09d}" for j in range(min(n, 20000))]
    codes = {"s_name": i % len(names), "s_address": uniform(62, i, 0, len(tpch.COMMENTS) - 1), "s_phone": uniform(63, i, 0, 24),
             "s_comment": uniform(64, i, 0, len(tpch.COMMENTS) - 1), "s_rev": torch.zeros_like(i)}
    vocab = {"s_name": names, "s_address": tpch.COMMENTS, "s_phone": [f"{10 + j}-{100 + j}-{200 + j}-{1000 + j}" for j in range(25)],
             "s_comment": tpch.COMMENTS, "s_rev": [""]}
    return RawTable("supplier", SUPPLIER_SCHEMA, n, cols, codes, vocab)


def gen_part(sf: float, device="cpu") -> RawTable:
    n = tpch.n_parts(sf)
    i = torch.arange(n, dtype=torch.int64, device=device)
    pk = i + 1
    cols = {"p_partkey": pk, "p_size": uniform(70, i, 1, 50), "p_retailprice": 90000 + (pk // 10) % 20001 + 100 * (pk % 1000)}
    mfgr = uniform(71, i, 1, 5)
    codes = {"p_name": uniform(72, i, 0, len(P_NAMES) - 1), "p_mfgr": mfgr - 1, "p_brand": (mfgr - 1) * 5 + uniform(73, i, 0, 4),
             "p_type": uniform(74, i, 0, len(P_TYPES) - 1), "p_container": uniform(75, i, 0, len(CONTAINERS) - 1),
             "p_comment": uniform(76, i, 0, len(tpch.COMMENTS) - 1), "p_rev": torch.zeros_like(i)}
    vocab = {"p_name": P_NAMES, "p_mfgr": [f"Manufacturer#{j}" for j in range(1, 6)],
             "p_brand": [f"Brand#{a}{b}" for a in range(1, 6) for b in range(1, 6)], "p_type": P_TYPES, "p_container": CONTAINERS,
             "p_comment": tpch.COMMENTS, "p_rev": [""]}
    return RawTable("part", PART_SCHEMA, n, cols, codes, vocab)


def gen_partsupp(sf: float, device="cpu") -> RawTable:
    """Four suppliers per part (tpch.supplier_of): lineitem's (l_partkey, l_suppkey) pairs are drawn from these rows."""
    n = tpch.n_parts(sf) * 4
    i = torch.arange(n, dtype=torch.int64, device=device)
    pk, j = i // 4 + 1, i % 4
    cols = {"ps_partkey": pk, "ps_suppkey": tpch.supplier_of(pk, j, sf), "ps_availqty": uniform(80, i, 1, 9999),
            "ps_supplycost": uniform(81, i, 100, 100000)}
    codes = {"ps_comment": uniform(82, i, 0, len(tpch.COMMENTS) - 1), "ps_rev": torch.zeros_like(i)}
    return RawTable("partsupp", PARTSUPP_SCHEMA, n, cols, codes, {"ps_comment": tpch.COMMENTS, "ps_rev": [""]})


def gen_nation(device="cpu") -> RawTable:
    i = torch.arange(25, dtype=torch.int64, device=device)
    cols = {"n_nationkey": i, "n_regionkey": torch.tensor([r for _, r in NATIONS], dtype=torch.int64, device=device)}
    codes = {"n_name": i.clone(), "n_comment": uniform(90, i, 0, len(tpch.COMMENTS) - 1), "n_rev": torch.zeros_like(i)}
    return RawTable("nation", NATION_SCHEMA, 25, cols, codes, {"n_name": [n for n, _ in NATIONS], "n_comment": tpch.COMMENTS, "n_rev": [""]})


def gen_region(device="cpu") -> RawTable:
    i = torch.arange(5, dtype=torch.int64, device=device)
    codes = {"r_name": i.clone(), "r_comment": uniform(91, i, 0, len(tpch.COMMENTS) - 1), "r_rev": torch.zeros_like(i)}
    return RawTable("region", REGION_SCHEMA, 5, {"r_regionkey": i}, codes, {"r_name": REGIONS, "r_comment": tpch.COMMENTS, "r_rev": [""]})


@dataclass
class FullDatabase:
    sf: float
    customer: MemoryTable
    orders: MemoryTable
    lineitem: MemoryTable
    supplier: MemoryTable
    part: MemoryTable
    partsupp: MemoryTable
    nation: MemoryTable
    region: MemoryTable


def generate_full(sf: float, batch_rows: Optional[int] = 1024) -> FullDatabase:
    """All eight tables as host MemoryTables (1024-row batches like the reference's CSV reader, csv.rs:63-66)."""
    raws = [tpch.gen_customer(sf), tpch.gen_orders(sf), tpch.gen_lineitem(sf), gen_supplier(sf), gen_part(sf), gen_partsupp(sf),
            gen_nation(), gen_region()]
    return FullDatabase(sf, *[MemoryTable.try_new(t.schema, tpch.to_arrow(t, batch_rows)) for t in raws])


# ------------------------------------------------------------------------------------------------
# expression helpers: result types and the reference's coercions
# ------------------------------------------------------------------------------------------------
def type_of(e: PhysicalExpr, schema: pa.Schema) -> pa.DataType:
    """LogicalExpr::data_type as far as these plans need it (decimal rules: arrow-rs add / sub / mul)."""
    if isinstance(e, Column):
        return schema.field(e.index).type
    if isinstance(e, Literal):
        return e.value.data_type
    if isinstance(e, CastExpr):
        return e.data_type
    if isinstance(e, (Like,)):
        return pa.bool_()
    if isinstance(e, Function):
        return pa.int64()
    if isinstance(e, CaseExpr):
        return type_of(e.when_then[0][1], schema)
    if isinstance(e, BinaryExpr):
        if e.op in (O.Eq, O.NotEq, O.Gt, O.GtEq, O.Lt, O.LtEq, O.And, O.Or):
            return pa.bool_()
        lt, rt = type_of(e.left, schema), type_of(e.right, schema)
        if pa.types.is_decimal(lt) and pa.types.is_decimal(rt):
            if e.op == O.Mul:
                return pa.decimal128(min(38, lt.precision + rt.precision + 1), lt.scale + rt.scale)
            if e.op in (O.Add, O.Sub):
                s = max(lt.scale, rt.scale)
                return pa.decimal128(min(38, max(lt.precision - lt.scale, rt.precision - rt.scale) + s + 1), s)
        if lt == rt:
            return lt
    raise TypeError(f"type_of: unsupported expression {e}")


def _cast_if(e: PhysicalExpr, have: pa.DataType, want: pa.DataType) -> PhysicalExpr:
    return e if have == want else CastExpr(e, want)


def binop(l: PhysicalExpr, op: Operator, r: PhysicalExpr, schema: pa.Schema) -> BinaryExpr:
    """BinaryExpr with the casts optimizer/rule/type_coercion.rs inserts from utils/type_coercion.rs:37-170."""
    if op in (O.And, O.Or):
        return BinaryExpr(l, op, r)
    lt, rt = type_of(l, schema), type_of(r, schema)
    is_int = pa.types.is_signed_integer
    if op in (O.Eq, O.NotEq, O.Gt, O.GtEq, O.Lt, O.LtEq):
        if lt == pa.date32() and pa.types.is_string(rt):
            return BinaryExpr(l, op, CastExpr(r, pa.date32()))
        if pa.types.is_string(lt) and rt == pa.date32():
            return BinaryExpr(CastExpr(l, pa.date32()), op, r)
        if pa.types.is_decimal(lt) and (is_int(rt) or pa.types.is_floating(rt)):
            return BinaryExpr(l, op, CastExpr(r, lt))
        if pa.types.is_decimal(rt) and (is_int(lt) or pa.types.is_floating(lt)):
            return BinaryExpr(CastExpr(l, rt), op, r)
        return BinaryExpr(l, op, r)
    if op == O.Div and (pa.types.is_decimal(lt) or pa.types.is_decimal(rt)):
        return BinaryExpr(_cast_if(l, lt, pa.float64()), op, _cast_if(r, rt, pa.float64()))
    int_dec = {pa.int8(): pa.decimal128(3, 0), pa.int16(): pa.decimal128(5, 0), pa.int32(): pa.decimal128(10, 0),
               pa.int64(): pa.decimal128(20, 0)}
    if pa.types.is_decimal(lt) and pa.types.is_decimal(rt):
        return BinaryExpr(l, op, r)
    if pa.types.is_decimal(lt) and rt in int_dec:
        return BinaryExpr(l, op, CastExpr(r, int_dec[rt]))
    if lt in int_dec and pa.types.is_decimal(rt):
        return BinaryExpr(CastExpr(l, int_dec[lt]), op, r)
    if lt == rt:
        return BinaryExpr(l, op, r)
    if pa.float64() in (lt, rt):   # numeric_coercion
        return BinaryExpr(_cast_if(l, lt, pa.float64()), op, _cast_if(r, rt, pa.float64()))
    if pa.int64() in (lt, rt):
        return BinaryExpr(_cast_if(l, lt, pa.int64()), op, _cast_if(r, rt, pa.int64()))
    raise TypeError(f"can not coerce type: {lt} and {rt} for numeric operation")


def conj(*es: PhysicalExpr) -> PhysicalExpr:
    out = es[0]
    for e in es[1:]:
        out = BinaryExpr(out, O.And, e)
    return out


def disj(*es: PhysicalExpr) -> PhysicalExpr:
    out = es[0]
    for e in es[1:]:
        out = BinaryExpr(out, O.Or, e)
    return out


def col(schema: pa.Schema, name: str, nth: int = 0) -> Column:
    """The nth field called `name` (a self-join carries the same name twice: nation n1 / nation n2)."""
    hits = [i for i, f in enumerate(schema) if f.name == name]
    return Column(name, hits[nth])


def utf8(s: str) -> Literal:
    return Literal(ScalarValue.Utf8(s))


def i64(v: int) -> Literal:
    return Literal(ScalarValue.Int64(v))


def date(s: str) -> CastExpr:
    return CastExpr(utf8(s), pa.date32())


def extract_year(e: PhysicalExpr) -> Function:
    return Function(DatetimeExtract(), [utf8("YEAR"), e])


def scan(t: MemoryTable, pred=None) -> Scan:
    return Scan(t.schema, t, None, pred(t.schema) if pred is not None else None)


def join(left, right, pairs: Sequence[Tuple[str, str]], join_type=JoinType.Inner, filter=None, nth=None) -> HashJoinExec:
    """HashJoinExec on (left column name, right column name) pairs; nth = {name: occurrence} for repeated names."""
    nth = nth or {}
    on = [(col(left.schema, a, nth.get(a, 0)), col(right.schema, b)) for a, b in pairs]
    return HashJoinExec.try_new(left, right, join_type, on, filter)


def volume(schema: pa.Schema) -> BinaryExpr:
    """l_extendedprice * (1 - l_discount): Decimal128(15,2) * (Decimal128(20,0) - Decimal128(15,2)) -> Decimal128(38,4)"""
    return binop(col(schema, "l_extendedprice"), O.Mul, binop(i64(1), O.Sub, col(schema, "l_discount"), schema), schema)


def sort_limit(plan, keys: Sequence[Tuple[PhysicalExpr, bool]], limit: Optional[int] = None):
    """ORDER BY (nulls_first = true, planner/mod.rs:339-342) [+ LIMIT n -> Limit(Sort.new_with_limit(n)), :69-83]."""
    exprs = [PhyscialSortExpr(e, SortOptions(desc, True)) for e, desc in keys]
    if limit is None:
        return Sort(exprs, plan)
    return Limit(Sort.new_with_limit(exprs, plan, limit), limit, 0)


def project(plan, items: Sequence[Tuple[str, PhysicalExpr]]) -> Projection:
    schema = pa.schema([(n, type_of(e, plan.schema)) for n, e in items])
    return Projection(schema, plan, [e for _, e in items])


def aggregate(plan, keys: Sequence[Tuple[str, PhysicalExpr]], aggs: Sequence[Tuple[str, object]]) -> HashAggregate:
    fields = [(n, type_of(e, plan.schema)) for n, e in keys] + [(n, a.return_type if hasattr(a, "return_type") else pa.int64()) for n, a in aggs]
    return HashAggregate(pa.schema(fields), plan, [e for _, e in keys], [a for _, a in aggs])


def sum_of(e: PhysicalExpr, schema: pa.Schema) -> SumAggregateExpr:
    return SumAggregateExpr(e, type_of(e, schema))       # AggregateOperator::infer_type: SUM keeps the argument type


# ------------------------------------------------------------------------------------------------
# the queries
# ------------------------------------------------------------------------------------------------
def q2_plan(db: FullDatabase):
    """tests/tpch/q2.slt:1-45.  part x supplier stays a CrossJoin (no equi-condition links them); the correlated
    `ps_supplycost = (select min(ps_supplycost) ... where p_partkey = ps_partkey ...)` is a LEFT join against the
    subquery grouped by its correlated column, then a Filter (scalar_subquery_to_join.rs:38-96)."""
    p = scan(db.part, lambda s: conj(binop(col(s, "p_size"), O.Eq, i64(15), s), Like(False, col(s, "p_type"), utf8("%BRASS"))))
    j = CrossJoin.new(p, scan(db.supplier))
    j = join(j, scan(db.partsupp), [("p_partkey", "ps_partkey"), ("s_suppkey", "ps_suppkey")])
    j = join(j, scan(db.nation), [("s_nationkey", "n_nationkey")])
    j = join(j, scan(db.region, lambda s: binop(col(s, "r_name"), O.Eq, utf8("EUROPE"), s)), [("n_regionkey", "r_regionkey")])
    # the subquery: partsupp, supplier, nation, region (EUROPE) grouped by ps_partkey
    sq = join(scan(db.partsupp), scan(db.supplier), [("ps_suppkey", "s_suppkey")])
    sq = join(sq, scan(db.nation), [("s_nationkey", "n_nationkey")])
    sq = join(sq, scan(db.region, lambda s: binop(col(s, "r_name"), O.Eq, utf8("EUROPE"), s)), [("n_regionkey", "r_regionkey")])
    sq = aggregate(sq, [("ps_partkey", col(sq.schema, "ps_partkey"))],
                   [("MIN(ps_supplycost)", MinAggregateExpr(col(sq.schema, "ps_supplycost"), DEC))])
    sq = project(sq, [("MIN(ps_supplycost)", Column("MIN(ps_supplycost)", 1)), ("ps_partkey", Column("ps_partkey", 0))])   # alias __scalar_sq_1
    lj = HashJoinExec.try_new(j, sq, JoinType.Left, [(col(j.schema, "p_partkey"), Column("ps_partkey", 1))], None)
    s = lj.schema
    f = Filter(lj, binop(col(s, "ps_supplycost"), O.Eq, col(s, "MIN(ps_supplycost)"), s))
    out = project(f, [(n, col(s, n)) for n in ("s_acctbal", "s_name", "n_name", "p_partkey", "p_mfgr", "s_address", "s_phone", "s_comment")])
    os_ = out.schema
    return sort_limit(out, [(col(os_, "s_acctbal"), True), (col(os_, "n_name"), False), (col(os_, "s_name"), False),
                            (col(os_, "p_partkey"), False)], 10)


def q4_plan(db: FullDatabase):
    """tests/tpch/q4.slt:1-22: EXISTS -> LeftSemi join orders x lineitem(l_commitdate < l_receiptdate)."""
    o = scan(db.orders, lambda s: conj(binop(col(s, "o_orderdate"), O.GtEq, utf8("1993-07-01"), s),
                                        binop(col(s, "o_orderdate"), O.Lt, date("1993-10-01"), s)))
    l = scan(db.lineitem, lambda s: binop(col(s, "l_commitdate"), O.Lt, col(s, "l_receiptdate"), s))
    sj = join(o, l, [("o_orderkey", "l_orderkey")], JoinType.LeftSemi)
    agg = aggregate(sj, [("o_orderpriority", col(sj.schema, "o_orderpriority"))], [("order_count", CountAggregateExpr(i64(1)))])
    out = project(agg, [("o_orderpriority", Column("o_orderpriority", 0)), ("order_count", Column("order_count", 1))])
    return sort_limit(out, [(Column("o_orderpriority", 0), False)])


def q5_plan(db: FullDatabase):
    """tests/tpch/q5.slt:1-24"""
    c = scan(db.customer)
    o = scan(db.orders, lambda s: conj(binop(col(s, "o_orderdate"), O.GtEq, date("1994-01-01"), s),
                                        binop(col(s, "o_orderdate"), O.Lt, date("1995-01-01"), s)))
    j = join(c, o, [("c_custkey", "o_custkey")])
    j = join(j, scan(db.lineitem), [("o_orderkey", "l_orderkey")])
    j = join(j, scan(db.supplier), [("l_suppkey", "s_suppkey"), ("c_nationkey", "s_nationkey")])
    j = join(j, scan(db.nation), [("s_nationkey", "n_nationkey")])
    j = join(j, scan(db.region, lambda s: binop(col(s, "r_name"), O.Eq, utf8("ASIA"), s)), [("n_regionkey", "r_regionkey")])
    s = j.schema
    agg = aggregate(j, [("n_name", col(s, "n_name"))], [("revenue", sum_of(volume(s), s))])
    out = project(agg, [("n_name", Column("n_name", 0)), ("revenue", Column("revenue", 1))])
    return sort_limit(out, [(Column("revenue", 1), True)])


def q7_plan(db: FullDatabase):
    """tests/tpch/q7.slt:1-39: the nation-pair predicate spans n1 and n2 and stays in a Filter above the joins."""
    l = scan(db.lineitem, lambda s: conj(binop(col(s, "l_shipdate"), O.GtEq, date("1995-01-01"), s),
                                          binop(col(s, "l_shipdate"), O.LtEq, date("1996-12-31"), s)))
    j = join(scan(db.supplier), l, [("s_suppkey", "l_suppkey")])
    j = join(j, scan(db.orders), [("l_orderkey", "o_orderkey")])
    j = join(j, scan(db.customer), [("o_custkey", "c_custkey")])
    j = join(j, scan(db.nation), [("s_nationkey", "n_nationkey")])             # n1
    j = join(j, scan(db.nation), [("c_nationkey", "n_nationkey")])             # n2
    s = j.schema
    n1, n2 = col(s, "n_name", 0), col(s, "n_name", 1)
    f = Filter(j, disj(conj(binop(n1, O.Eq, utf8("FRANCE"), s), binop(n2, O.Eq, utf8("GERMANY"), s)),
                       conj(binop(n1, O.Eq, utf8("GERMANY"), s), binop(n2, O.Eq, utf8("FRANCE"), s))))
    shipping = project(f, [("supp_nation", n1), ("cust_nation", n2), ("l_year", extract_year(col(s, "l_shipdate"))), ("volume", volume(s))])
    ss = shipping.schema
    agg = aggregate(shipping, [(n, col(ss, n)) for n in ("supp_nation", "cust_nation", "l_year")], [("revenue", sum_of(col(ss, "volume"), ss))])
    out = project(agg, [(f.name, Column(f.name, i)) for i, f in enumerate(agg.schema)])
    return sort_limit(out, [(Column("supp_nation", 0), False), (Column("cust_nation", 1), False), (Column("l_year", 2), False)])


def q8_plan(db: FullDatabase):
    """tests/tpch/q8.slt:1-37: part x supplier is a CrossJoin; decimal `/` runs in Float64 (type_coercion.rs:109-115)."""
    p = scan(db.part, lambda s: binop(col(s, "p_type"), O.Eq, utf8("ECONOMY ANODIZED STEEL"), s))
    j = CrossJoin.new(p, scan(db.supplier))
    j = join(j, scan(db.lineitem), [("p_partkey", "l_partkey"), ("s_suppkey", "l_suppkey")])
    o = scan(db.orders, lambda s: conj(binop(col(s, "o_orderdate"), O.GtEq, date("1995-01-01"), s),
                                        binop(col(s, "o_orderdate"), O.LtEq, date("1996-12-31"), s)))
    j = join(j, o, [("l_orderkey", "o_orderkey")])
    j = join(j, scan(db.customer), [("o_custkey", "c_custkey")])
    j = join(j, scan(db.nation), [("c_nationkey", "n_nationkey")])             # n1
    j = join(j, scan(db.nation), [("s_nationkey", "n_nationkey")])             # n2
    j = join(j, scan(db.region, lambda s: binop(col(s, "r_name"), O.Eq, utf8("AMERICA"), s)), [("n_regionkey", "r_regionkey")])  # n1.n_regionkey
    s = j.schema
    all_nations = project(j, [("o_year", extract_year(col(s, "o_orderdate"))), ("volume", volume(s)), ("nation", col(s, "n_name", 1))])
    a = all_nations.schema
    vt = type_of(col(a, "volume"), a)
    brazil = CaseExpr([(binop(col(a, "nation"), O.Eq, utf8("BRAZIL"), a), col(a, "volume"))], CastExpr(i64(0), vt))
    agg = aggregate(all_nations, [("o_year", col(a, "o_year"))], [("sum_case", SumAggregateExpr(brazil, vt)), ("sum_volume", sum_of(col(a, "volume"), a))])
    g = agg.schema
    d122 = pa.decimal128(12, 2)
    share = CastExpr(binop(CastExpr(col(g, "sum_case"), d122), O.Div, CastExpr(col(g, "sum_volume"), d122), g), DEC)
    out = project(agg, [("o_year", col(g, "o_year")), ("mkt_share", share)])
    return sort_limit(out, [(Column("o_year", 0), False)])


def q9_plan(db: FullDatabase):
    """tests/tpch/q9.slt:1-33"""
    p = scan(db.part, lambda s: Like(False, col(s, "p_name"), utf8("%green%")))
    j = CrossJoin.new(p, scan(db.supplier))
    j = join(j, scan(db.lineitem), [("s_suppkey", "l_suppkey"), ("p_partkey", "l_partkey")])
    j = join(j, scan(db.partsupp), [("l_suppkey", "ps_suppkey"), ("l_partkey", "ps_partkey")])
    j = join(j, scan(db.orders), [("l_orderkey", "o_orderkey")])
    j = join(j, scan(db.nation), [("s_nationkey", "n_nationkey")])
    s = j.schema
    amount = binop(volume(s), O.Sub, binop(col(s, "ps_supplycost"), O.Mul, col(s, "l_quantity"), s), s)
    profit = project(j, [("nation", col(s, "n_name")), ("o_year", extract_year(col(s, "o_orderdate"))), ("amount", amount)])
    ps = profit.schema
    agg = aggregate(profit, [("nation", col(ps, "nation")), ("o_year", col(ps, "o_year"))], [("sum_profit", sum_of(col(ps, "amount"), ps))])
    out = project(agg, [(f.name, Column(f.name, i)) for i, f in enumerate(agg.schema)])
    return sort_limit(out, [(Column("nation", 0), False), (Column("o_year", 1), True)], 10)


def q10_plan(db: FullDatabase):
    """tests/tpch/q10.slt:1-31: seven group keys, five of them strings."""
    o = scan(db.orders, lambda s: conj(binop(col(s, "o_orderdate"), O.GtEq, date("1993-10-01"), s),
                                        binop(col(s, "o_orderdate"), O.Lt, date("1994-01-01"), s)))
    j = join(scan(db.customer), o, [("c_custkey", "o_custkey")])
    j = join(j, scan(db.lineitem, lambda s: binop(col(s, "l_returnflag"), O.Eq, utf8("R"), s)), [("o_orderkey", "l_orderkey")])
    j = join(j, scan(db.nation), [("c_nationkey", "n_nationkey")])
    s = j.schema
    keys = ["c_custkey", "c_name", "c_acctbal", "c_phone", "n_name", "c_address", "c_comment"]
    agg = aggregate(j, [(k, col(s, k)) for k in keys], [("revenue", sum_of(volume(s), s))])
    g = agg.schema
    out = project(agg, [(n, col(g, n)) for n in ("c_custkey", "c_name", "revenue", "c_acctbal", "n_name", "c_address", "c_phone", "c_comment")])
    return sort_limit(out, [(Column("revenue", 2), True)], 10)


def q11_plan(db: FullDatabase):
    """tests/tpch/q11.slt:1-27: the uncorrelated HAVING subquery is a LEFT nested-loop join with the filter `true`
    (scalar_subquery_to_join.rs:86-88); `sum * 0.0001` runs in Float64 (numeric_coercion) and the comparison casts it back
    to the sum's decimal type (type_coercion.rs:76-85)."""
    def chain():
        j = join(scan(db.partsupp), scan(db.supplier), [("ps_suppkey", "s_suppkey")])
        return join(j, scan(db.nation, lambda s: binop(col(s, "n_name"), O.Eq, utf8("GERMANY"), s)), [("s_nationkey", "n_nationkey")])
    j = chain()
    s = j.schema
    value = binop(col(s, "ps_supplycost"), O.Mul, col(s, "ps_availqty"), s)
    agg = aggregate(j, [("ps_partkey", col(s, "ps_partkey"))], [("value", sum_of(value, s))])
    from .physical.plan import NoGroupingAggregate
    j2 = chain()
    total = NoGroupingAggregate(pa.schema([("total", type_of(value, s))]), j2, [sum_of(value, j2.schema)])
    ts = total.schema
    sq = project(total, [("threshold", binop(col(ts, "total"), O.Mul, Literal(ScalarValue.Float64(0.0001)), ts))])     # alias __scalar_sq_1
    true_filter = JoinFilter(Literal(ScalarValue.Boolean(True)), pa.schema([]), [])
    lj = NestedLoopJoinExec.try_new(agg, sq, JoinType.Left, true_filter)
    ls = lj.schema
    f = Filter(lj, binop(col(ls, "value"), O.Gt, col(ls, "threshold"), ls))
    out = project(f, [("ps_partkey", col(ls, "ps_partkey")), ("value", col(ls, "value"))])
    return sort_limit(out, [(Column("value", 1), True)], 10)


def q12_plan(db: FullDatabase):
    """tests/tpch/q12.slt:1-29: explicit JOIN; IN-list as an OR chain; CASE -> Int64 sums."""
    l = scan(db.lineitem, lambda s: conj(disj(binop(col(s, "l_shipmode"), O.Eq, utf8("MAIL"), s), binop(col(s, "l_shipmode"), O.Eq, utf8("SHIP"), s)),
                                          binop(col(s, "l_commitdate"), O.Lt, col(s, "l_receiptdate"), s),
                                          binop(col(s, "l_shipdate"), O.Lt, col(s, "l_commitdate"), s),
                                          binop(col(s, "l_receiptdate"), O.GtEq, date("1994-01-01"), s),
                                          binop(col(s, "l_receiptdate"), O.Lt, date("1995-01-01"), s)))
    j = join(l, scan(db.orders), [("l_orderkey", "o_orderkey")])
    s = j.schema
    pr = col(s, "o_orderpriority")
    high = CaseExpr([(disj(binop(pr, O.Eq, utf8("1-URGENT"), s), binop(pr, O.Eq, utf8("2-HIGH"), s)), i64(1))], i64(0))
    low = CaseExpr([(conj(binop(pr, O.NotEq, utf8("1-URGENT"), s), binop(pr, O.NotEq, utf8("2-HIGH"), s)), i64(1))], i64(0))
    agg = aggregate(j, [("l_shipmode", col(s, "l_shipmode"))],
                    [("high_line_count", SumAggregateExpr(high, pa.int64())), ("low_line_count", SumAggregateExpr(low, pa.int64()))])
    out = project(agg, [(f.name, Column(f.name, i)) for i, f in enumerate(agg.schema)])
    return sort_limit(out, [(Column("l_shipmode", 0), False)])


QUERIES = {"q2": q2_plan, "q4": q4_plan, "q5": q5_plan, "q7": q7_plan, "q8": q8_plan, "q9": q9_plan, "q10": q10_plan,
           "q11": q11_plan, "q12": q12_plan}
