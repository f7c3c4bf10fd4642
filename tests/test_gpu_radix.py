"""Radix-partitioned aggregate (csrc/radix_agg.cuh) -- the high-cardinality group-by path of BASELINE.json configs[3].
Forced at small sizes with QGPU_RADIX=force and compared with the oracle (HashAggregate, hash.rs:45-107,138-170):
bit-exact keys / COUNT / integer and decimal SUM / MIN / MAX, Float64 SUM/AVG within 1e-12 relative."""
import decimal

import numpy as np
import pyarrow as pa
import pytest

from oracle import qref
from qurious_b200.physical.expr import (AvgAggregateExpr, Column, CountAggregateExpr, MaxAggregateExpr, MinAggregateExpr,
                                        SumAggregateExpr)
from qurious_b200.physical.plan import HashAggregate, MemoryTable, Scan
from tests.cases import bx, check_rows, lit, rows_of

pytestmark = pytest.mark.gpu

DEC = pa.decimal128(15, 2)


def table(n, keys, seed=1, extremes=False, skew=False):
    rng = np.random.default_rng(seed)
    k = rng.integers(0, keys, n).astype(np.int64)
    with np.errstate(over="ignore"):
        k = (k.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)).view(np.int64)
    if skew:
        k[rng.random(n) < 0.5] = 42
    if extremes and n >= 4:
        k[:4] = [np.iinfo(np.int64).min, np.iinfo(np.int64).max, np.iinfo(np.int64).max, 0]
    d = rng.integers(8000, 8000 + 37, n).astype(np.int32)
    v = rng.integers(-10**6, 10**6, n).astype(np.int64)
    f = rng.random(n)
    p = [decimal.Decimal(int(x)).scaleb(-2) for x in rng.integers(-10**9, 10**9, n)]
    schema = pa.schema([("k", pa.int64()), ("d", pa.date32()), ("v", pa.int64()), ("f", pa.float64()), ("p", DEC)])
    cols = [pa.array(k), pa.array(d, pa.date32()), pa.array(v), pa.array(f), pa.array(p, DEC)]
    return MemoryTable.try_new(schema, [pa.record_batch(cols, schema=schema)])


def plan_of(t, keys=("k",), predicate=None):
    s = t.schema
    C = lambda n: Column(n, s.get_field_index(n))  # noqa: E731
    aggs = [SumAggregateExpr(C("v"), pa.int64()), CountAggregateExpr(C("v")), MinAggregateExpr(C("v"), pa.int64()),
            MaxAggregateExpr(C("v"), pa.int64()), AvgAggregateExpr(C("f"), pa.float64(), pa.float64()),
            SumAggregateExpr(C("p"), pa.decimal128(25, 2))]
    out = pa.schema([(n, s.field(n).type) for n in keys] + [(f"a{i}", a.return_type) for i, a in enumerate(aggs)])
    return HashAggregate(out, Scan(s, t, None, predicate), [C(n) for n in keys], aggs)


@pytest.mark.parametrize("n,keys,kw", [(200_000, 50_000, {}), (300_000, 3_000, {"skew": True}), (70_000, 70_000, {"extremes": True}),
                                        (4097, 4097, {}), (5, 3, {})])
def test_radix_groupby_matches_oracle(gpu_ctx, monkeypatch, n, keys, kw):
    monkeypatch.setenv("QGPU_RADIX", "force")
    t = table(n, keys, seed=n % 97, **kw)
    p = plan_of(t)
    got = rows_of(p.execute(gpu_ctx))
    if not kw.get("skew"):      # a heavy hitter overflows its bucket's staging area: the query re-runs on the HBM hash table
        assert "radix-partitioned" in p.last_strategy(), p.last_strategy()
    check_rows("radix", got, rows_of(qref.execute(plan_of(t))), ordered=False)
    again = rows_of(p.execute(gpu_ctx))          # cached plan, second execution
    check_rows("radix again", again, got, ordered=False)


def test_radix_two_keys_predicate_and_decimal_key(gpu_ctx, monkeypatch):
    monkeypatch.setenv("QGPU_RADIX", "force")
    t = table(150_000, 900, seed=5)
    s = t.schema
    pred = bx(Column("v", s.get_field_index("v")), "Gt", lit(-500_000))
    p = plan_of(t, keys=("d", "v"), predicate=pred)
    got = rows_of(p.execute(gpu_ctx))
    assert "radix-partitioned" in p.last_strategy(), p.last_strategy()
    check_rows("radix 2 keys", got, rows_of(qref.execute(plan_of(t, keys=("d", "v"), predicate=pred))), ordered=False)
    p2 = plan_of(t, keys=("p",))
    got2 = rows_of(p2.execute(gpu_ctx))
    assert "radix-partitioned" in p2.last_strategy(), p2.last_strategy()
    check_rows("radix decimal key", got2, rows_of(qref.execute(plan_of(t, keys=("p",)))), ordered=False)


def test_radix_overflow_falls_back_to_hash(gpu_ctx, monkeypatch):
    monkeypatch.setenv("QGPU_RADIX", "force")
    monkeypatch.setenv("QGPU_RADIX_TEST_CAP", "64")     # 512 buckets x 64 slots < 60k groups
    t = table(120_000, 60_000, seed=11)
    p = plan_of(t)
    got = rows_of(p.execute(gpu_ctx))
    assert "hbm-hash" in p.last_strategy(), p.last_strategy()
    check_rows("fallback", got, rows_of(qref.execute(plan_of(t))), ordered=False)


def test_radix_off_and_default_small_inputs_use_hash(gpu_ctx, monkeypatch):
    monkeypatch.delenv("QGPU_RADIX", raising=False)
    t = table(50_000, 20_000, seed=2)
    p = plan_of(t)
    got = rows_of(p.execute(gpu_ctx))
    assert "hbm-hash" in p.last_strategy(), p.last_strategy()
    check_rows("hash", got, rows_of(qref.execute(plan_of(t))), ordered=False)


def test_radix_predicate_that_keeps_nothing_or_little(gpu_ctx, monkeypatch):
    monkeypatch.setenv("QGPU_RADIX", "force")
    t = table(50_000, 9_000, seed=8)
    s = t.schema
    for bound in (10**7, 999_990):            # no row / a handful of rows pass
        pred = bx(Column("v", s.get_field_index("v")), "Gt", lit(bound))
        p = plan_of(t, predicate=pred)
        check_rows(f"radix v > {bound}", rows_of(p.execute(gpu_ctx)), rows_of(qref.execute(plan_of(t, predicate=pred))), ordered=False)


# ---- the TMA-pipelined scatter / histogram kernels and the 16-byte pair layout (tuples of <= 2 operand values) -------------
def narrow_table(n, keys, seed, key_type=pa.int64(), v_type=pa.int64()):
    rng = np.random.default_rng(seed)
    k = rng.integers(0, keys, n).astype(np.int64)
    if key_type == pa.int64():
        with np.errstate(over="ignore"):
            k = (k.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)).view(np.int64)
        if n >= 3:
            k[:3] = [-1, np.iinfo(np.int64).max, np.iinfo(np.int64).min]   # -1 - min(k) wraps to the EMPTY marker's code
    lim = 30_000
    v = rng.integers(-lim, lim, n).astype(np.int64)
    g = rng.integers(0, 5, n).astype(np.uint8)
    f = rng.random(n)
    schema = pa.schema([("k", key_type), ("g", pa.uint8()), ("v", v_type), ("f", pa.float64()), ("w", pa.int64())])
    cols = [pa.array(k.astype(key_type.to_pandas_dtype())), pa.array(g), pa.array(v.astype(v_type.to_pandas_dtype())), pa.array(f), pa.array(v)]
    return MemoryTable.try_new(schema, [pa.record_batch(cols, schema=schema)])


def narrow_plan(t, shape, keys=("k",), predicate=None):
    s = t.schema
    C = lambda n: Column(n, s.get_field_index(n))  # noqa: E731
    vt = s.field("v").type
    sum_v = [SumAggregateExpr(C("v"), vt)] if vt == pa.int64() else []   # SUM exists for 64-bit integers only (sum.rs:44-46)
    aggs = {"vf": sum_v + [CountAggregateExpr(C("v")), MinAggregateExpr(C("v"), vt),
                           MaxAggregateExpr(C("v"), vt), AvgAggregateExpr(C("f"), pa.float64(), pa.float64())],
            "v": sum_v + [MaxAggregateExpr(C("v"), vt), MinAggregateExpr(C("v"), vt)],
            "f": [SumAggregateExpr(C("f"), pa.float64()), AvgAggregateExpr(C("f"), pa.float64(), pa.float64()), CountAggregateExpr(C("f"))],
            "count": [CountAggregateExpr(C("v"))]}[shape]
    out = pa.schema([(n, s.field(n).type) for n in keys] + [(f"a{i}", a.return_type) for i, a in enumerate(aggs)])
    return HashAggregate(out, Scan(s, t, None, predicate), [C(n) for n in keys], aggs)


@pytest.mark.parametrize("pair", ["1", "0"])
@pytest.mark.parametrize("n,keys", [(5, 3), (4096, 1500), (4097, 4097), (100_003, 31_000), (300_000, 20_000)])
def test_radix_tma_scatter_two_values(gpu_ctx, monkeypatch, n, keys, pair):
    """The shape of BASELINE.json configs[3] (SUM / COUNT / MIN / MAX(v), AVG(f)): TMA scatter on both levels, the two operand
    values as 16-byte pairs (pair == "1") or as separate arrays."""
    monkeypatch.setenv("QGPU_RADIX", "force")
    monkeypatch.setenv("QGPU_RADIX_PAIR", pair)
    t = narrow_table(n, keys, seed=n % 89 + 1)
    p = narrow_plan(t, "vf")
    got = rows_of(p.execute(gpu_ctx))
    assert "radix-partitioned" in p.last_strategy() and "result columns written by the final pass" in p.last_strategy(), p.last_strategy()
    assert "scatter tma/tma" in p.last_strategy() and ("16 B value pairs" in p.last_strategy()) == (pair == "1"), p.last_strategy()
    check_rows(f"radix tma pair={pair}", got, rows_of(qref.execute(narrow_plan(t, "vf"))), ordered=False)
    check_rows("radix tma again", rows_of(p.execute(gpu_ctx)), got, ordered=False)


@pytest.mark.parametrize("shape", ["v", "f", "count"])
@pytest.mark.parametrize("key_type,v_type", [(pa.int64(), pa.int64()), (pa.int32(), pa.int16()), (pa.int32(), pa.int32())])
def test_radix_tma_scatter_shapes_and_widths(gpu_ctx, monkeypatch, shape, key_type, v_type):
    """One / no operand value per tuple, narrow key and value columns in the staged tiles, a second (uint8) key and a range
    predicate evaluated from the stage."""
    monkeypatch.setenv("QGPU_RADIX", "force")
    t = narrow_table(130_001, 9_000, seed=17, key_type=key_type, v_type=v_type)
    s = t.schema
    for keys, pred in ((("k",), None), (("k", "g"), bx(Column("w", s.get_field_index("w")), "Gt", lit(-20_000)))):
        p = narrow_plan(t, shape, keys=keys, predicate=pred)
        got = rows_of(p.execute(gpu_ctx))
        if len(keys) == 1 or key_type != pa.int64():   # (a full-range int64 key + a second key do not pack into 64 bits)
            assert "radix-partitioned" in p.last_strategy() and "scatter tma/tma" in p.last_strategy(), p.last_strategy()
        check_rows(f"radix tma {shape} {keys}", got, rows_of(qref.execute(narrow_plan(t, shape, keys=keys, predicate=pred))), ordered=False)


def test_radix_register_staged_scatter_still_matches(gpu_ctx, monkeypatch):
    monkeypatch.setenv("QGPU_RADIX", "force")
    monkeypatch.setenv("QGPU_RADIX_SCATTER", "regs")
    t = narrow_table(90_000, 11_000, seed=3)
    p = narrow_plan(t, "vf")
    got = rows_of(p.execute(gpu_ctx))
    assert "scatter regs/regs" in p.last_strategy(), p.last_strategy()
    check_rows("radix regs", got, rows_of(qref.execute(narrow_plan(t, "vf"))), ordered=False)


@pytest.mark.parametrize("shape", ["vf", "v", "count"])
def test_radix_sorting_form_of_the_final_pass_still_matches(gpu_ctx, monkeypatch, shape):
    """QGPU_RADIX_AGG=sort: the final pass that ranks, scans and re-stages the rows (the list form is the default)."""
    monkeypatch.setenv("QGPU_RADIX", "force")
    monkeypatch.setenv("QGPU_RADIX_AGG", "sort")
    t = narrow_table(120_000, 14_000, seed=29)
    p = narrow_plan(t, shape)
    got = rows_of(p.execute(gpu_ctx))
    assert "radix-partitioned" in p.last_strategy(), p.last_strategy()
    check_rows(f"radix sort form {shape}", got, rows_of(qref.execute(narrow_plan(t, shape))), ordered=False)
