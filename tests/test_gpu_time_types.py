"""Time32(Second | Millisecond) / Time64(Microsecond | Nanosecond): key types the reference's hasher accepts
(utils/array.rs:198-201) and its MIN / MAX accumulators instantiate (aggregate/mod.rs:104-107, then panic in
ScalarValue::try_from_array like the date types, scalar.rs:228).  They cross the C ABI with their unit, round-trip
through a table, serve as group keys, join keys, sort keys and comparison operands -- GPU against the oracle."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import qref
from qurious_b200 import QuriousError, _lib
from qurious_b200.datatypes import JoinType, Operator
from qurious_b200.physical.expr import BinaryExpr, Column, CountAggregateExpr, MinAggregateExpr, SumAggregateExpr
from qurious_b200.physical.plan import (HashAggregate, HashJoinExec, MemoryTable, PhyscialSortExpr, Projection, Scan, Sort,
                                        SortOptions)
from tests.cases import check_rows, rows_of

pytestmark = pytest.mark.gpu
TYPES = [pa.time32("s"), pa.time32("ms"), pa.time64("us"), pa.time64("ns")]


def table(dt, n=3000, seed=1, distinct=40):
    rng = np.random.default_rng(seed)
    hi = 86400 if dt == pa.time32("s") else 86_400_000
    base = pa.int32() if pa.types.is_time32(dt) else pa.int64()
    t = [None if rng.random() < 0.1 else int(v) * (hi // distinct) for v in rng.integers(0, distinct, n)]
    u = [None if rng.random() < 0.1 else int(v) * (hi // distinct) for v in rng.integers(0, distinct, n)]
    schema = pa.schema([("t", dt), ("u", dt), ("v", pa.int64())])
    b = pa.record_batch([pa.array(t, base).cast(dt), pa.array(u, base).cast(dt), pa.array(rng.integers(-50, 50, n))], schema=schema)
    return MemoryTable.try_new(schema, [b.slice(0, n // 3), b.slice(n // 3)])


@pytest.mark.parametrize("dt", TYPES, ids=str)
def test_round_trip_group_sort_compare(gpu_ctx, dt):
    t = table(dt)
    dev = t.device_table(gpu_ctx)
    assert dev.to_batch().schema.types == t.schema.types
    assert pa.Table.from_batches([dev.to_batch()]).equals(pa.Table.from_batches(t.data).combine_chunks())
    T, U, V = Column("t", 0), Column("u", 1), Column("v", 2)

    def agg():
        return HashAggregate(pa.schema([("t", dt), ("n", pa.int64()), ("s", pa.int64())]), Scan(t.schema, t, None, BinaryExpr(T, Operator.LtEq, U)),
                             [T], [CountAggregateExpr(V), SumAggregateExpr(V, pa.int64())])
    check_rows("group by time", rows_of(agg().execute(gpu_ctx)), rows_of(qref.execute(agg())), ordered=False)

    def srt():
        p = Projection(pa.schema([("t", dt), ("v", pa.int64())]), Scan(t.schema, t, None, None), [T, V])
        return Sort([PhyscialSortExpr(Column("t", 0), SortOptions(True, False)), PhyscialSortExpr(Column("v", 1), SortOptions(False, True))], p)
    check_rows("sort by time", rows_of(srt().execute(gpu_ctx)), rows_of(qref.execute(srt())), ordered=True)


@pytest.mark.parametrize("dt", [pa.time32("ms"), pa.time64("ns")], ids=str)
def test_time_join_key(gpu_ctx, dt):
    a, b = table(dt, 400, seed=2), table(dt, 500, seed=3)

    def make(jt):
        return HashJoinExec.try_new(Scan(a.schema, a, None, None), Scan(b.schema, b, None, None), jt, [(Column("t", 0), Column("u", 1))], None)
    for jt in (JoinType.Inner, JoinType.Left, JoinType.LeftAnti):
        check_rows(f"join {jt.name}", rows_of(make(jt).execute(gpu_ctx)), rows_of(qref.execute(make(jt))), ordered=True)


def test_min_of_time_is_unimplemented_like_the_reference(gpu_ctx):
    dt = pa.time64("us")
    t = table(dt, 100)
    p = HashAggregate(pa.schema([("v", pa.int64()), ("m", dt)]), Scan(t.schema, t, None, None), [Column("v", 2)], [MinAggregateExpr(Column("t", 0), dt)])
    with pytest.raises(QuriousError):
        p.execute(gpu_ctx)
    with pytest.raises(qref.QError):
        qref.execute(p)
