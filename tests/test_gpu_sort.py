"""Sort / Limit on the GPU (csrc/sort.cu; SURVEY 8f "next" #1) against the oracle's restatement of sort.rs:48-82 and
limit.rs:27-58: same rows in the same ORDER (stable for equal keys), every key type, NULL placement, DESC, top-N."""
import decimal

import numpy as np
import pyarrow as pa
import pytest

from oracle import qref
from qurious_b200 import tpch
from qurious_b200.physical.expr import Column
from qurious_b200.physical.plan import Limit, MemoryTable, PhyscialSortExpr, Scan, Sort, SortOptions
from tests.cases import bx, check_rows, lit, rows_of

pytestmark = pytest.mark.gpu


def nullable(rng, vals, t, frac=0.12):
    m = rng.random(len(vals)) < frac
    return pa.array([None if d else v for v, d in zip(vals, m)], type=t)


def mixed_table(n, seed, batches=1):
    rng = np.random.default_rng(seed)
    f = rng.normal(0, 10, n).round(1)
    if n:
        f[rng.integers(0, n, max(n // 50, 1))] = np.nan
        f[rng.integers(0, n, max(n // 50, 1))] = -0.0
        f[rng.integers(0, n, max(n // 50, 1))] = np.inf
    words = ["", "a", "ab", "ab\x00", "abc", "b", "BUILDING", "zz", "a much longer string value", "é"]
    cols = {
        "i": nullable(rng, rng.integers(-5, 6, n).tolist(), pa.int64()),
        "f": nullable(rng, f.tolist(), pa.float64()),
        "s": nullable(rng, [words[k] for k in rng.integers(0, len(words), n)], pa.string()),
        "d": nullable(rng, rng.integers(9000, 9010, n).astype(np.int32).tolist(), pa.date32()),
        "w": pa.array([None if x % 13 == 0 else decimal.Decimal(int(x)).scaleb(-2) for x in rng.integers(-10**17, 10**17, n)] if n else [],
                      pa.decimal128(30, 2)),
        "p": pa.array([decimal.Decimal(int(x)).scaleb(-2) for x in rng.integers(-500, 500, n)], pa.decimal128(15, 2)),
        "b": nullable(rng, (rng.random(n) < 0.5).tolist(), pa.bool_()),
        "u": pa.array(rng.integers(0, 2**63, n).astype(np.uint64) * 2, pa.uint64()),
        "id": pa.array(np.arange(n), pa.int64()),
    }
    schema = pa.schema([(k, v.type) for k, v in cols.items()])
    full = pa.record_batch(list(cols.values()), schema=schema)
    if batches == 1 or n == 0:
        return MemoryTable.try_new(schema, [full])
    step = max(n // batches, 1)
    return MemoryTable.try_new(schema, [full.slice(o, min(step, n - o)) for o in range(0, n, step)])


def col(t, name):
    return Column(name, t.schema.get_field_index(name))


@pytest.mark.parametrize("keys,limit", [
    ([("i", False, True)], None), ([("i", True, False)], 7), ([("f", False, True)], None), ([("f", True, False)], None),
    ([("s", False, True)], None), ([("s", True, True), ("i", False, False)], None), ([("d", True, True), ("b", False, False)], 100),
    ([("w", False, False)], None), ([("w", True, True)], 33), ([("p", False, True), ("f", True, True)], None),
    ([("u", True, True)], None), ([("b", False, True), ("i", True, True), ("s", False, False)], None), ([], None), ([("id", True, True)], 0)])
def test_sort_matches_oracle_in_order(gpu_ctx, keys, limit):
    for n, batches in ((0, 1), (1, 1), (37, 1), (5000, 3)):
        t = mixed_table(n, seed=n + len(keys), batches=batches)

        def make():
            return Sort([PhyscialSortExpr(col(t, k), SortOptions(descending=d, nulls_first=nf)) for k, d, nf in keys],
                        Scan(t.schema, t, None, None), limit)
        got = make().execute(gpu_ctx)
        assert len(got) <= 1
        check_rows(f"sort {keys} n={n}", rows_of(got), rows_of(qref.execute(make())), ordered=True)


def test_sort_on_expression_and_after_filter(gpu_ctx):
    t = mixed_table(3000, seed=5)

    def make():
        s = Scan(t.schema, t, None, bx(col(t, "id"), "Gt", lit(100)))
        return Limit(Sort([PhyscialSortExpr(bx(col(t, "i"), "Mul", col(t, "i")), SortOptions(True, False)),
                           PhyscialSortExpr(col(t, "s"), SortOptions(False, True))], s), 50, 5)
    check_rows("sort expr + limit", rows_of(make().execute(gpu_ctx)), rows_of(qref.execute(make())), ordered=True)


@pytest.mark.parametrize("fetch,skip", [(3, 0), (None, 2), (2, 2), (10**6, 0), (0, 0), (None, 10**6), (1, 36)])
def test_limit_matches_oracle(gpu_ctx, fetch, skip):
    t = mixed_table(37, seed=2)

    def make():
        return Limit(Scan(t.schema, t, None, None), fetch, skip)
    check_rows(f"limit {fetch} {skip}", rows_of(make().execute(gpu_ctx)), rows_of(qref.execute(make())), ordered=True)


def test_tpch_q1_q3_with_order_by_and_limit(gpu_ctx):
    db = tpch.generate(0.02, batch_rows=None)
    for make in (tpch.q1_sorted_plan, tpch.q3_top10_plan):
        p = make(db)
        got = rows_of(p.execute(gpu_ctx))
        assert "sort[" in p.last_strategy(), p.last_strategy()
        check_rows(make.__name__, got, rows_of(qref.execute(make(db))), ordered=True)
    assert len(rows_of(tpch.q3_top10_plan(db).execute(gpu_ctx))) == 10
