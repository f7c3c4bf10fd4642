"""Reference quirks that only exist behind a switch: by default the library returns the mathematically
intended result; with qgpu_set_compat it reproduces the reference's failure -- as the oracle's compat=True mode restates it
(avg.rs:105-116: the pre-division value is validated against the target precision and a NULL Decimal128(38,10) comes back,
which fails the RecordBatch schema check; sum.rs:101: an all-NULL / empty decimal SUM is a NULL typed Decimal128(38,10))."""
import decimal

import pyarrow as pa
import pytest

from oracle import qref
from qurious_b200 import QuriousError
from qurious_b200.physical.expr import AvgAggregateExpr, Column, CountAggregateExpr, SumAggregateExpr, avg_return_type
from qurious_b200.physical.plan import HashAggregate, MemoryTable, NoGroupingAggregate, Scan
from tests.cases import bx, check_rows, lit, rows_of

pytestmark = pytest.mark.gpu
DEC = pa.decimal128(15, 2)


@pytest.fixture
def compat(gpu_ctx):
    def on(name):
        gpu_ctx.set_compat(name, True)
    yield on
    gpu_ctx.set_compat("avg_precision", False)
    gpu_ctx.set_compat("empty_decimal_sum", False)


def big_avg_table():
    # sum = 2 * 9.9e12 (raw 1.98e15); x 10^4 = 1.98e19 > 10^19 - 1: does not fit Decimal128(19, 6)'s precision
    v = decimal.Decimal("9900000000000.00")
    schema = pa.schema([("k", pa.int64()), ("p", DEC)])
    return MemoryTable.try_new(schema, [pa.record_batch([pa.array([1, 1, 2]), pa.array([v, v, decimal.Decimal("1.50")], DEC)], schema=schema)])


def avg_plan(t, grouped):
    P = Column("p", 1)
    rt = avg_return_type(DEC)
    if grouped:
        return HashAggregate(pa.schema([("k", pa.int64()), ("a", rt)]), Scan(t.schema, t, None, None), [Column("k", 0)],
                             [AvgAggregateExpr(P, DEC, rt)])
    return NoGroupingAggregate(pa.schema([("a", rt)]), Scan(t.schema, t, None, None), [AvgAggregateExpr(P, DEC, rt)])


@pytest.mark.parametrize("grouped", [False, True])
def test_decimal_avg_precision_quirk(gpu_ctx, compat, grouped):
    t = big_avg_table()
    # default: sum * 10^4 / count, truncated (what SF100 Q1 needs)
    check_rows("avg default", rows_of(avg_plan(t, grouped).execute(gpu_ctx)), rows_of(qref.execute(avg_plan(t, grouped))), ordered=False)
    # compat: the reference's behaviour is an error
    with pytest.raises(qref.QError):
        qref.execute(avg_plan(t, grouped), compat=True)
    compat("avg_precision")
    with pytest.raises(QuriousError):
        avg_plan(t, grouped).execute(gpu_ctx)


def test_empty_decimal_sum_quirk(gpu_ctx, compat):
    schema = pa.schema([("k", pa.int64()), ("p", DEC)])
    t = MemoryTable.try_new(schema, [pa.record_batch([pa.array([1, 2, 3]), pa.array([decimal.Decimal("1.00")] * 3, DEC)], schema=schema)])

    def plan():
        return NoGroupingAggregate(pa.schema([("s", DEC), ("c", pa.int64())]), Scan(schema, t, None, bx(Column("k", 0), "Gt", lit(100))),
                                   [SumAggregateExpr(Column("p", 1), DEC), CountAggregateExpr(Column("p", 1))])
    got = rows_of(plan().execute(gpu_ctx))
    assert got == [(None, 0)]                      # aggregation.slt:162-170,192-195: SUM over no rows is NULL, COUNT is 0
    check_rows("empty sum default", got, rows_of(qref.execute(plan())), ordered=True)
    with pytest.raises(qref.QError):
        qref.execute(plan(), compat=True)
    compat("empty_decimal_sum")
    with pytest.raises(QuriousError):
        plan().execute(gpu_ctx)


def test_unknown_compat_switch_is_an_error(gpu_ctx):
    with pytest.raises(QuriousError):
        gpu_ctx.set_compat("no_such_switch", True)
