"""String-valued expressions on the GPU against the oracle: CASE over Utf8 branches (physical/expr/case.rs:30-47 folds
`zip(mask, then, acc)` over string arrays), Utf8 literals as projected columns (literal.rs:19-23), and their use as
filter operands and group keys."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import qref
from qurious_b200.datatypes import Operator, ScalarValue
from qurious_b200.physical.expr import BinaryExpr, CaseExpr, Column, CountAggregateExpr, IsNull, Literal
from qurious_b200.physical.plan import Filter, HashAggregate, MemoryTable, Projection, Scan
from tests.cases import check_rows, rows_of

pytestmark = pytest.mark.gpu
O = Operator
WORDS = ["", "a", "MAIL", "SHIP", "1-URGENT", "forest green", "héllo wörld", "x" * 70]


def table(n, seed, batches=1):
    rng = np.random.default_rng(seed)
    s1 = [None if rng.random() < 0.15 else WORDS[k] for k in rng.integers(0, len(WORDS), n)]
    s2 = [None if rng.random() < 0.15 else WORDS[k] + "!" for k in rng.integers(0, len(WORDS), n)]
    x = [None if rng.random() < 0.1 else int(v) for v in rng.integers(-10, 10, n)]
    schema = pa.schema([("s1", pa.string()), ("s2", pa.string()), ("x", pa.int64())])
    b = pa.record_batch([pa.array(s1, pa.string()), pa.array(s2, pa.string()), pa.array(x, pa.int64())], schema=schema)
    step = max(1, n // batches)
    return MemoryTable.try_new(schema, [b.slice(o, step) for o in range(0, n, step)] if n else [b])


def utf8(v):
    return Literal(ScalarValue.Utf8(v))


def i64(v):
    return Literal(ScalarValue.Int64(v))


S1, S2, X = Column("s1", 0), Column("s2", 1), Column("x", 2)
CASE3 = CaseExpr([(BinaryExpr(X, O.Gt, i64(0)), S1), (BinaryExpr(X, O.Lt, i64(-5)), utf8("lit-branch"))], S2)
CASE_NULL_ELSE = CaseExpr([(IsNull(S1), utf8("was null"))], Literal(ScalarValue.Utf8(None)))


@pytest.mark.parametrize("n,batches", [(0, 1), (1, 1), (31, 1), (33, 2), (5000, 7)])
def test_string_case_projection(gpu_ctx, n, batches):
    t = table(n, seed=n + 1, batches=batches)

    def make():
        return Projection(pa.schema([("c", pa.string()), ("k", pa.string()), ("e", pa.string()), ("x", pa.int64())]),
                          Scan(t.schema, t, None, None), [CASE3, utf8("constant"), CASE_NULL_ELSE, X])
    check_rows("string case", rows_of(make().execute(gpu_ctx)), rows_of(qref.execute(make())), ordered=True)


def test_string_case_as_filter_operand_and_group_key(gpu_ctx):
    t = table(4000, seed=3, batches=3)

    def filt():
        return Filter(Scan(t.schema, t, None, None), BinaryExpr(CASE3, O.Eq, utf8("MAIL")))
    check_rows("filter on case", rows_of(filt().execute(gpu_ctx)), rows_of(qref.execute(filt())), ordered=True)

    def agg():
        proj = Projection(pa.schema([("c", pa.string()), ("x", pa.int64())]), Scan(t.schema, t, None, None), [CASE3, X])
        return HashAggregate(pa.schema([("c", pa.string()), ("n", pa.int64())]), proj, [Column("c", 0)], [CountAggregateExpr(Column("x", 1))])
    check_rows("group by case", rows_of(agg().execute(gpu_ctx)), rows_of(qref.execute(agg())), ordered=False)


def test_string_case_branch_type_mismatch_is_an_arrow_error(gpu_ctx):
    from qurious_b200 import QuriousError
    t = table(10, seed=1)
    bad = CaseExpr([(BinaryExpr(X, O.Gt, i64(0)), S1)], i64(0))
    p = Projection(pa.schema([("c", pa.string())]), Scan(t.schema, t, None, None), [bad])
    with pytest.raises(QuriousError) as e:
        p.execute(gpu_ctx)
    assert e.value.kind == "ArrowError"
    with pytest.raises(qref.QError):
        qref.execute(p)
