"""Like / NotLike (physical/expr/like.rs:28-41 -> arrow like/nlike) and EXTRACT (functions/datetime/extract.rs) on the GPU
against the oracle: in projections, in filters and as group keys (TPC-H Q7-Q9 shapes); SURVEY 8f "next" #3."""
import datetime

import numpy as np
import pyarrow as pa
import pytest

from oracle import qref
from qurious_b200 import QuriousError
from qurious_b200.datatypes import ScalarValue
from qurious_b200.physical.expr import (Column, CountAggregateExpr, DatetimeExtract, Function, Like, Literal, SumAggregateExpr)
from qurious_b200.physical.plan import HashAggregate, MemoryTable, Projection, Scan
from tests.cases import check_rows, rows_of

pytestmark = pytest.mark.gpu

WORDS = ["", "a", "ab", "a%b", "a_b", "axb", "BUILDING", "forest green", "green forest", "special requests", "héllo wörld", "%", "_",
         "a\\b", "ends with %", "xx special packages requests xx"]
PATTERNS = ["%", "", "_", "a%", "%b", "a_b", "a\\%b", "a\\_b", "%green%", "forest%", "%special%requests%", "h_llo%", "%%", "_%_", "a\\\\b",
            "BUILDING", "%\\%", "%wörld", "héllo w_rld"]


def table(n, seed=1):
    rng = np.random.default_rng(seed)
    s = [None if rng.random() < 0.1 else WORDS[k] for k in rng.integers(0, len(WORDS), n)]
    p = [None if rng.random() < 0.1 else PATTERNS[k] for k in rng.integers(0, len(PATTERNS), n)]
    d = [None if rng.random() < 0.1 else datetime.date(1970, 1, 1) + datetime.timedelta(days=int(x)) for x in rng.integers(-40000, 40000, n)]
    ms = [None if x is None else (x - datetime.date(1970, 1, 1)).days * 86400000 for x in d]
    schema = pa.schema([("s", pa.string()), ("p", pa.string()), ("d", pa.date32()), ("d64", pa.date64()), ("v", pa.int64())])
    b = pa.record_batch([pa.array(s, pa.string()), pa.array(p, pa.string()), pa.array(d, pa.date32()), pa.array(ms, pa.date64()),
                         pa.array(rng.integers(-100, 100, n))], schema=schema)
    return MemoryTable.try_new(schema, [b])


def utf8(x):
    return Literal(ScalarValue.Utf8(x))


@pytest.mark.parametrize("pattern", PATTERNS)
def test_like_literal_pattern(gpu_ctx, pattern):
    t = table(600, seed=len(pattern))
    S = Column("s", 0)

    def make():
        return Projection(pa.schema([("l", pa.bool_()), ("n", pa.bool_())]), Scan(t.schema, t, None, None),
                          [Like(False, S, utf8(pattern)), Like(True, S, utf8(pattern))])
    check_rows(f"like {pattern!r}", rows_of(make().execute(gpu_ctx)), rows_of(qref.execute(make())), ordered=True)


def test_like_column_pattern_and_filter(gpu_ctx):
    t = table(3000, seed=9)
    S, P, V = Column("s", 0), Column("p", 1), Column("v", 4)

    def make():
        return Projection(pa.schema([("s", pa.string()), ("p", pa.string()), ("m", pa.bool_())]),
                          Scan(t.schema, t, None, Like(False, S, utf8("%e%"))), [S, P, Like(False, S, P)])
    check_rows("like col", rows_of(make().execute(gpu_ctx)), rows_of(qref.execute(make())), ordered=True)

    def agg():
        return HashAggregate(pa.schema([("s", pa.string()), ("c", pa.int64()), ("sv", pa.int64())]),
                             Scan(t.schema, t, None, Like(True, S, utf8("%special%requests%"))), [S],
                             [CountAggregateExpr(V), SumAggregateExpr(V, pa.int64())])
    check_rows("nlike filter + agg", rows_of(agg().execute(gpu_ctx)), rows_of(qref.execute(agg())), ordered=False)


def test_like_needs_strings(gpu_ctx):
    t = table(10)
    p = Projection(pa.schema([("m", pa.bool_())]), Scan(t.schema, t, None, None), [Like(False, Column("v", 4), utf8("%"))])
    with pytest.raises(QuriousError):
        p.execute(gpu_ctx)
    with pytest.raises(qref.QError):
        qref.execute(p)


def test_extract_parts_projection_and_group_key(gpu_ctx):
    t = table(5000, seed=3)
    D, D64, V = Column("d", 2), Column("d64", 3), Column("v", 4)

    def ex(part, c):
        return Function(DatetimeExtract(), [utf8(part), c])

    def make():
        return Projection(pa.schema([(n, pa.int64()) for n in ("y", "m", "dd", "y64", "m64", "d64")]), Scan(t.schema, t, None, None),
                          [ex("YEAR", D), ex("month", D), ex("Day", D), ex("year", D64), ex("MONTH", D64), ex("day", D64)])
    check_rows("extract", rows_of(make().execute(gpu_ctx)), rows_of(qref.execute(make())), ordered=True)

    def agg():  # TPC-H Q7/Q8/Q9: extract(year from date) as a group key
        return HashAggregate(pa.schema([("y", pa.int64()), ("c", pa.int64()), ("sv", pa.int64())]), Scan(t.schema, t, None, None),
                             [ex("year", D)], [CountAggregateExpr(V), SumAggregateExpr(V, pa.int64())])
    check_rows("group by extract(year)", rows_of(agg().execute(gpu_ctx)), rows_of(qref.execute(agg())), ordered=False)


def test_extract_unsupported_part_is_an_error(gpu_ctx):
    t = table(10)
    p = Projection(pa.schema([("c", pa.int64())]), Scan(t.schema, t, None, None),
                   [Function(DatetimeExtract(), [utf8("century"), Column("d", 2)])])
    with pytest.raises(QuriousError):
        p.execute(gpu_ctx)
    with pytest.raises(qref.QError):
        qref.execute(p)
