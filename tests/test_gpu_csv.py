"""Device CSV / .tbl parser (csrc/csv.cu; SURVEY 8f #4: file -> HBM ingest, replacing read_csv / COPY ... FROM of
datasource/file/csv.rs:16-72 + planner/sql.rs:324-375 for tables with a declared schema).  Checked against the values the text
was written from, against Arrow C++'s CSV reader (pyarrow.csv with the same column types), and end to end: TPC-H Q1 / Q6 / Q3
over tables loaded from `|`-delimited .tbl text (trailing delimiter -> the NULL `*_rev` columns of create_tables.slt) equal
the plans over the Arrow-ingested tables."""
import io
import os
import tempfile

import numpy as np
import pyarrow as pa
import pyarrow.csv as pacsv
import pytest

from qurious_b200 import QuriousError, _lib, tpch
from qurious_b200.physical.plan import MemoryTable
from tests.cases import check_rows, rows_of

pytestmark = pytest.mark.gpu


def tbl_text(mt: MemoryTable) -> bytes:
    """dbgen's .tbl layout: fields joined by '|', every line ends with '|'; the last (`*_rev`) column is that empty field."""
    t = pa.Table.from_batches(mt.data)
    cols = []
    for name in t.schema.names[:-1]:
        c = t[name]
        if pa.types.is_decimal(c.type):
            cols.append([str(v) for v in c.to_pylist()])
        elif c.type == pa.date32():
            cols.append([v.isoformat() for v in c.to_pylist()])
        else:
            cols.append([str(v) for v in c.to_pylist()])
    return ("\n".join("|".join(r) + "|" for r in zip(*cols)) + "\n").encode()


def load(ctx, schema, source, **kw) -> pa.Table:
    dev = _lib.DeviceTable.create(ctx, schema)
    try:
        n = dev.append_csv(source, **kw)
        assert n == dev.num_rows
        return pa.Table.from_batches([dev.to_batch()])
    finally:
        dev.free()


def test_tpch_tbl_files_round_trip_and_queries(gpu_ctx):
    db = tpch.generate(0.01, batch_rows=None)
    loaded = {}
    with tempfile.TemporaryDirectory() as d:
        for name in ("customer", "orders", "lineitem"):
            mt = getattr(db, name)
            path = os.path.join(d, name + ".tbl")
            with open(path, "wb") as f:
                f.write(tbl_text(mt))
            dev = _lib.DeviceTable.create(gpu_ctx, mt.schema)
            rows = dev.append_csv(path, has_header=False, delimiter="|")          # COPY name FROM 'name.tbl' (DELIMITER '|')
            want = pa.Table.from_batches(mt.data)
            assert rows == want.num_rows
            got = pa.Table.from_batches([dev.to_batch()])
            rev = want.schema.names[-1]
            assert got[rev].null_count == rows                                    # the trailing '|' -> an empty field -> NULL
            assert got.drop([rev]).equals(want.drop([rev]).combine_chunks())
            loaded[name] = MemoryTable.from_device_table(dev)
    db_csv = tpch.Database(0.01, loaded["customer"], loaded["orders"], loaded["lineitem"])
    for q in ("q6", "q1", "q3"):
        a = getattr(tpch, q + "_plan")(db_csv).execute(gpu_ctx)
        b = getattr(tpch, q + "_plan")(db).execute(gpu_ctx)
        check_rows(q + " over .tbl", rows_of(a), rows_of(b), ordered=(q != "q1"))
    for t in loaded.values():
        t._dev.free()


SCHEMA = pa.schema([("i", pa.int64()), ("s", pa.string()), ("f", pa.float64()), ("d", pa.date32()), ("m", pa.decimal128(15, 2)),
                    ("w", pa.decimal128(30, 4)), ("b", pa.bool_()), ("k", pa.int32()), ("u", pa.uint8())])


def random_csv(n, seed, newline="\n", trailing_newline=True, header=True):
    rng = np.random.default_rng(seed)

    def maybe(x, p=0.1):
        return "" if rng.random() < p else x
    lines = [",".join(SCHEMA.names)] if header else []
    words = ["alpha", "b", "c d", "é", "x" * 40, "-", "0"]
    for _ in range(n):
        lines.append(",".join([
            maybe(str(int(rng.integers(-2**62, 2**62)))), maybe(words[int(rng.integers(0, len(words)))]),
            maybe(rng.choice(["%.6f" % rng.standard_normal(), "%d" % rng.integers(-999, 999), "%.3e" % (rng.standard_normal() * 1e5), "1.5E-7", "-0.0"])),
            maybe("%04d-%02d-%02d" % (rng.integers(1900, 2100), rng.integers(1, 13), rng.integers(1, 29))),
            maybe(rng.choice(["%d.%02d" % (rng.integers(-10**9, 10**9), rng.integers(0, 100)), "%d" % rng.integers(-5, 5), "%d.5" % rng.integers(0, 9), "-.25"])),
            maybe("%d.%04d" % (rng.integers(0, 10**15), rng.integers(0, 10**4))), maybe(rng.choice(["true", "false", "TRUE", "False"])),
            maybe(str(int(rng.integers(-2**31, 2**31)))), maybe(str(int(rng.integers(0, 256))))]))
    return (newline.join(lines) + (newline if trailing_newline else "")).encode()


@pytest.mark.parametrize("n,newline,trailing,header", [(0, "\n", True, True), (1, "\n", False, False), (33, "\r\n", True, True),
                                                       (5000, "\n", True, True), (4097, "\n", False, True)])
def test_against_arrow_cpp_csv_reader(gpu_ctx, n, newline, trailing, header):
    text = random_csv(n, seed=n + 1, newline=newline, trailing_newline=trailing, header=header)
    got = load(gpu_ctx, SCHEMA, text, has_header=header, delimiter=",")
    ref = pacsv.read_csv(io.BytesIO(text), read_options=pacsv.ReadOptions(column_names=None if header else SCHEMA.names),
                         convert_options=pacsv.ConvertOptions(column_types=SCHEMA, strings_can_be_null=True, null_values=[""],
                                                              true_values=["true", "TRUE", "True"], false_values=["false", "FALSE", "False"])) \
        if (n > 0 or header) else SCHEMA.empty_table()
    assert got.num_rows == n
    assert got.schema.types == SCHEMA.types
    for name in SCHEMA.names:
        assert got[name].to_pylist() == ref[name].to_pylist(), name


def test_errors(gpu_ctx):
    two = pa.schema([("a", pa.int64()), ("m", pa.decimal128(15, 2))])
    with pytest.raises(QuriousError) as e:
        load(gpu_ctx, two, b"1,2.50\n3\n5,6.00\n", has_header=False)
    assert e.value.kind == "ArrowError" and "line 2" in str(e.value)
    with pytest.raises(QuriousError) as e:
        load(gpu_ctx, two, b"1,2.50\nx,1.00\n", has_header=False)
    assert e.value.kind == "ArrowError" and "line 2" in str(e.value) and "column 1" in str(e.value)
    with pytest.raises(QuriousError) as e:
        load(gpu_ctx, two, b"1,2.505\n", has_header=False)                       # more fractional digits than the scale
    assert e.value.kind == "InternalError"
    with pytest.raises(QuriousError) as e:
        load(gpu_ctx, two, b'1,"2.50"\n', has_header=False, quote='"')
    assert e.value.kind == "InternalError" and "quot" in str(e.value)
    with pytest.raises(QuriousError):
        load(gpu_ctx, two, "/nonexistent/file.tbl", has_header=False)
