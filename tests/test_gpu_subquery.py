"""`SubQuery` (physical/expr/subquery.rs:11-20) on the GPU against the oracle: the sub-plan's first column stands in as an
array operand; a scalar subquery in a SELECT list over a one-row relation works, a length mismatch is the ArrowError arrow's
binary kernels raise."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import qref
from qurious_b200 import QuriousError
from qurious_b200.datatypes import Operator, ScalarValue
from qurious_b200.physical.expr import BinaryExpr, Column, Literal, MaxAggregateExpr, SubQuery, SumAggregateExpr
from qurious_b200.physical.plan import MemoryTable, NoGroupingAggregate, Projection, Scan
from tests.cases import check_rows, rows_of

pytestmark = pytest.mark.gpu
O = Operator


def table(n, seed=1):
    rng = np.random.default_rng(seed)
    schema = pa.schema([("v", pa.int64()), ("w", pa.int64())])
    return MemoryTable.try_new(schema, [pa.record_batch([pa.array(rng.integers(-100, 100, n)), pa.array(rng.integers(0, 50, n))], schema=schema)])


def test_scalar_subquery_in_a_select_list(gpu_ctx):
    """SELECT (SELECT max(v) FROM t) + sum(w) FROM t  -- both sides are one-row relations"""
    t = table(5000)

    def make():
        sub = NoGroupingAggregate(pa.schema([("MAX(v)", pa.int64())]), Scan(t.schema, t, None, None), [MaxAggregateExpr(Column("v", 0), pa.int64())])
        outer = NoGroupingAggregate(pa.schema([("SUM(w)", pa.int64())]), Scan(t.schema, t, None, None), [SumAggregateExpr(Column("w", 1), pa.int64())])
        return Projection(pa.schema([("r", pa.int64()), ("m", pa.int64())]), outer,
                          [BinaryExpr(SubQuery(sub), O.Add, Column("SUM(w)", 0)), SubQuery(sub)])
    check_rows("scalar subquery", rows_of(make().execute(gpu_ctx)), rows_of(qref.execute(make())), ordered=True)


def test_subquery_column_of_the_same_length_and_in_a_filter(gpu_ctx):
    t = table(3000, seed=2)

    def make():
        sub = Projection(pa.schema([("x", pa.int64())]), Scan(t.schema, t, None, None), [BinaryExpr(Column("w", 1), O.Mul, Literal(ScalarValue.Int64(2)))])
        return Projection(pa.schema([("v", pa.int64()), ("x", pa.int64())]),
                          Scan(t.schema, t, None, BinaryExpr(SubQuery(sub), O.Gt, Column("v", 0))), [Column("v", 0), SubQuery(sub)])
    # the Scan's pushed-down filter sees the whole batch (3000 rows = the sub-plan's rows); the Projection above sees fewer rows
    # than the sub-plan returns: an ArrowError in the reference's kernels, the oracle and the library alike
    with pytest.raises(QuriousError) as e:
        make().execute(gpu_ctx)
    assert e.value.kind == "ArrowError"
    with pytest.raises(qref.QError):
        qref.execute(make())

    def filt_only():
        sub = Projection(pa.schema([("x", pa.int64())]), Scan(t.schema, t, None, None), [BinaryExpr(Column("w", 1), O.Mul, Literal(ScalarValue.Int64(2)))])
        return Scan(t.schema, t, None, BinaryExpr(SubQuery(sub), O.Gt, Column("v", 0)))
    check_rows("subquery in a filter", rows_of(filt_only().execute(gpu_ctx)), rows_of(qref.execute(filt_only())), ordered=True)
