"""Staged ingest (csrc/ingest.cu): many host RecordBatches -> one device chunk, at the reference's batch granularity
(MemoryTable::insert, datasource/memory.rs:104-111; 1024-row batches from datasource/file/csv.rs:34-72).  Every case
uploads through the C ABI and downloads again: the round trip must be value-identical to the concatenated input, whatever
the batch boundaries, slice offsets, NULL patterns, thread counts or the host-narrowing switch."""
import decimal
import random

import numpy as np
import pyarrow as pa
import pytest

from qurious_b200 import _lib
from qurious_b200.physical.plan import MemoryTable

pytestmark = pytest.mark.gpu


def make_table(n, seed, with_nulls=True, wide_decimal=False):
    rng = np.random.default_rng(seed)
    rnd = random.Random(seed)

    def nullify(vals, p=0.1):
        if not with_nulls:
            return vals
        return [None if rnd.random() < p else v for v in vals]
    i64 = nullify([int(x) for x in rng.integers(-2**62, 2**62, n)])
    i32 = nullify([int(x) for x in rng.integers(-2**31, 2**31 - 1, n)])
    d32 = nullify([int(x) for x in rng.integers(0, 20000, n)])
    f64 = nullify([float(x) for x in rng.standard_normal(n)])
    u8 = nullify([int(x) for x in rng.integers(0, 255, n)])
    bo = nullify([bool(x) for x in rng.integers(0, 2, n)])
    dec = nullify([decimal.Decimal(int(x)).scaleb(-2) for x in rng.integers(-10**14, 10**14, n)])
    big = 10**30 if wide_decimal else 10**14
    dec38 = nullify([decimal.Decimal(rnd.randint(-big, big)).scaleb(-4) for _ in range(n)])
    words = ["", "a", "BUILDING", "MACHINERY", "x" * 37, "déjà vu", "N", "O"]
    st = nullify([words[int(x)] for x in rng.integers(0, len(words), n)])
    return pa.table({
        "i64": pa.array(i64, pa.int64()), "i32": pa.array(i32, pa.int32()), "d32": pa.array(d32, pa.int32()).cast(pa.date32()),
        "f64": pa.array(f64, pa.float64()), "u8": pa.array(u8, pa.uint8()), "b": pa.array(bo, pa.bool_()),
        "dec": pa.array(dec, pa.decimal128(15, 2)), "dec38": pa.array(dec38, pa.decimal128(38, 4)),
        "s": pa.array(st, pa.utf8()), "nul": pa.nulls(n)})


def split(table, sizes):
    out, at = [], 0
    for s in sizes:
        out.extend(table.slice(at, s).to_batches() or [pa.RecordBatch.from_pylist([], schema=table.schema)])
        at += s
    assert at == table.num_rows
    return out


def round_trip(ctx, schema, batches, upload_columns=None, stream=True):
    dev = _lib.DeviceTable.create(ctx, schema)
    try:
        if stream:
            assert dev.append_batches(batches, upload_columns) == len(batches)
        else:
            for b in batches:
                dev.append(b, upload_columns)
        assert dev.num_rows == sum(b.num_rows for b in batches)
        return dev.to_batch()
    finally:
        dev.free()


def sizes_for(n, kind, seed):
    rnd = random.Random(seed)
    if kind == "one":
        return [n]
    if kind == "1024":
        return [1024] * (n // 1024) + ([n % 1024] if n % 1024 else [])
    out, left = [], n
    while left:
        s = min(left, rnd.choice([0, 1, 3, 7, 8, 31, 33, 64, 100, 1000, 4097]))
        out.append(s)
        left -= s
    return out


@pytest.mark.parametrize("kind", ["one", "1024", "ragged"])
@pytest.mark.parametrize("nulls", [False, True])
def test_round_trip_all_types(kind, nulls):
    ctx = _lib.default_context()
    t = make_table(20_000, 7, with_nulls=nulls)
    got = round_trip(ctx, t.schema, split(t, sizes_for(t.num_rows, kind, 3)))
    assert pa.Table.from_batches([got]).equals(t.combine_chunks())


def test_per_batch_append_equals_stream_and_sliced_children():
    ctx = _lib.default_context()
    t = make_table(9_000, 11)
    # batches whose child arrays carry non-zero offsets (slices of a bigger batch at odd positions)
    whole = t.combine_chunks().to_batches()[0]
    cuts = [0, 5, 1029, 1030, 4099, 9000]
    batches = [whole.slice(a, b - a) for a, b in zip(cuts, cuts[1:])]
    a = round_trip(ctx, t.schema, batches, stream=True)
    b = round_trip(ctx, t.schema, batches, stream=False)
    assert a.equals(b)
    assert pa.Table.from_batches([a]).equals(t.combine_chunks())


@pytest.mark.parametrize("threads,narrow", [(1, 1), (4, 1), (16, 0), (3, -1)])
def test_threads_and_host_narrow(threads, narrow):
    ctx = _lib.default_context()
    ctx.set_option("ingest_threads", threads)
    ctx.set_option("ingest_host_narrow", narrow)
    try:
        t = make_table(700_000, 5, with_nulls=(threads != 4))     # > 4 MiB: the worker threads run
        got = round_trip(ctx, t.schema, split(t, sizes_for(t.num_rows, "1024", 0)))
        assert pa.Table.from_batches([got]).equals(t.combine_chunks())
    finally:
        ctx.set_option("ingest_threads", 0)
        ctx.set_option("ingest_host_narrow", -1)


def test_decimal_that_does_not_fit_int64_stays_wide():
    """Decimal128(18, 0) declares <= 18 digits, but Arrow does not validate values: one value beyond int64 must send the
    whole column back as 16-byte values (the narrowing proof fails), not truncate it."""
    ctx = _lib.default_context()
    vals = [decimal.Decimal(i) for i in range(5000)]
    arr = pa.array(vals, pa.decimal128(38, 0))
    huge = pa.array([decimal.Decimal(2**70)], pa.decimal128(38, 0))
    col = pa.concat_arrays([arr, huge, arr])
    # re-label as Decimal128(18, 0) without validation
    col18 = pa.Array.from_buffers(pa.decimal128(18, 0), len(col), col.buffers())
    t = pa.table({"d": col18})
    got = round_trip(ctx, t.schema, split(t, [3000, 3000, 4001]))
    assert got.column(0).buffers()[1].to_pybytes()[:16 * len(col)] == col.buffers()[1].to_pybytes()[:16 * len(col)]


def test_upload_subset_and_query_over_small_batches():
    """Q1 / Q6 / Q3 over tables appended as 1024-row batches equal the same plans over one batch per table."""
    from qurious_b200 import tpch
    from tests.cases import check_rows, rows_of
    ctx = _lib.default_context()
    db1 = tpch.generate(0.01, batch_rows=None)
    dbk = tpch.generate(0.01)
    assert len(dbk.lineitem.data) > 50
    for q in ("q6", "q1", "q3"):
        a = getattr(tpch, q + "_plan")(db1).execute(ctx)
        b = getattr(tpch, q + "_plan")(dbk).execute(ctx)
        check_rows(q, rows_of(b), rows_of(a), ordered=(q != "q1"))


def test_append_after_use_and_mixed_upload_sets():
    ctx = _lib.default_context()
    t = make_table(3000, 2)
    bs = split(t, [1000, 1000, 1000])
    dev = _lib.DeviceTable.create(ctx, t.schema)
    try:
        dev.append(bs[0])
        dev.flush()
        dev.append_batches(bs[1:])
        assert pa.Table.from_batches([dev.to_batch()]).equals(t.combine_chunks())
    finally:
        dev.free()
    # a failing batch is reported by the append itself and leaves the table untouched
    dev = _lib.DeviceTable.create(ctx, t.schema)
    try:
        dev.append(bs[0])
        bad = bs[1].select([0, 1])
        with pytest.raises(_lib.QuriousError) as e:
            dev.append(bad)
        assert e.value.kind == "ArrowError"
        dev.append(bs[1])
        assert dev.num_rows == 2000
        assert pa.Table.from_batches([dev.to_batch()]).equals(t.slice(0, 2000).combine_chunks())
    finally:
        dev.free()


def test_empty_batches_only():
    ctx = _lib.default_context()
    t = make_table(0, 1)
    empty = pa.RecordBatch.from_pylist([], schema=t.schema)
    got = round_trip(ctx, t.schema, [empty, empty])
    assert got.num_rows == 0
