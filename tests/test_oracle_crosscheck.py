"""Independent cross-check of the oracle's TPC-H results (the reference's own TPC-H goldens cannot be used
-- dbgen data is not shipped -- so TPC-H *values* stay unpinned against the reference).  What CAN be pinned is the
oracle's arithmetic against the SQL definition of the queries: Q6 / Q1 / Q3 are recomputed here row by row with Python's
exact `decimal.Decimal` from the generated Arrow batches -- no numpy, none of the oracle's code -- and must equal the
oracle's output digit for digit (decimal AVG = sum * 10^4 / count truncated: avg.rs:105-116).  CPU only."""
import collections
import datetime
from decimal import ROUND_DOWN, Decimal

from oracle import qref
from qurious_b200 import tpch
from tests.cases import rows_of

SF = 0.002
D = datetime.date


def _rows(table):
    out = []
    for b in table.data:
        out.extend(b.to_pylist())
    return out


def _trunc(x: Decimal, digits: int) -> Decimal:
    return x.quantize(Decimal(1).scaleb(-digits), rounding=ROUND_DOWN)


def test_q6_matches_plain_decimal_arithmetic():
    db = tpch.generate(SF)
    total = Decimal(0)
    for r in _rows(db.lineitem):
        if (D(1994, 1, 1) <= r["l_shipdate"] < D(1995, 1, 1) and Decimal("0.05") <= r["l_discount"] <= Decimal("0.07")
                and r["l_quantity"] < 24):
            total += r["l_extendedprice"] * r["l_discount"]
    got = rows_of(qref.execute(tpch.q6_plan(db)))
    assert got == [(total,)]
    assert total != 0


def test_q1_matches_plain_decimal_arithmetic():
    db = tpch.generate(SF)
    acc = collections.defaultdict(lambda: [Decimal(0)] * 5 + [0])
    for r in _rows(db.lineitem):
        if r["l_shipdate"] > D(1998, 9, 2):
            continue
        a = acc[(r["l_returnflag"], r["l_linestatus"])]
        disc_price = r["l_extendedprice"] * (1 - r["l_discount"])
        a[0] += r["l_quantity"]
        a[1] += r["l_extendedprice"]
        a[2] += disc_price
        a[3] += disc_price * (1 + r["l_tax"])
        a[4] += r["l_discount"]
        a[5] += 1
    exp = sorted((k[0], k[1], a[0], a[1], a[2], a[3], _trunc(a[0] / a[5], 6), _trunc(a[1] / a[5], 6), _trunc(a[4] / a[5], 6), a[5])
                 for k, a in acc.items())
    got = sorted(rows_of(qref.execute(tpch.q1_plan(db))))
    assert got == exp
    assert len(exp) >= 3


def test_q3_matches_plain_decimal_arithmetic():
    db = tpch.generate(SF)
    building = {r["c_custkey"] for r in _rows(db.customer) if r["c_mktsegment"] == "BUILDING"}
    orders = {r["o_orderkey"]: r for r in _rows(db.orders) if r["o_orderdate"] < D(1995, 3, 15) and r["o_custkey"] in building}
    rev = collections.defaultdict(Decimal)
    for r in _rows(db.lineitem):
        o = orders.get(r["l_orderkey"])
        if o is not None and r["l_shipdate"] > D(1995, 3, 15):
            rev[(r["l_orderkey"], o["o_orderdate"], o["o_shippriority"])] += r["l_extendedprice"] * (1 - r["l_discount"])
    exp = sorted((k[0], v, k[1], k[2]) for k, v in rev.items())
    got = sorted(rows_of(qref.execute(tpch.q3_plan(db))))
    assert got == exp
    assert len(exp) > 0


def test_q6_matches_arrow_cpp():
    """SURVEY 8c "independent cross-check available here": Q6 through Arrow C++ (pyarrow compute) -- its decimal
    comparison / multiply / sum kernels -- must give the oracle's Decimal128(31,4) value exactly."""
    import pyarrow as pa
    import pyarrow.compute as pc
    db = tpch.generate(0.01)
    t = pa.Table.from_batches(db.lineitem.data)
    dec = pa.decimal128(15, 2)
    m = pc.and_(pc.and_(pc.greater_equal(t["l_shipdate"], D(1994, 1, 1)), pc.less(t["l_shipdate"], D(1995, 1, 1))),
                pc.and_(pc.and_(pc.greater_equal(t["l_discount"], pa.scalar(Decimal("0.05"), dec)),
                                pc.less_equal(t["l_discount"], pa.scalar(Decimal("0.07"), dec))),
                        pc.less(t["l_quantity"], pa.scalar(Decimal("24.00"), dec))))
    f = t.filter(m)
    exp = pc.sum(pc.multiply(f["l_extendedprice"], f["l_discount"])).as_py()
    assert rows_of(qref.execute(tpch.q6_plan(db))) == [(exp,)]
