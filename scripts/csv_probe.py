"""File -> HBM ingest probe: lineitem (Q1's columns) at the given SF written as '|'-delimited text by Arrow C++'s CSV writer,
then read + parsed on the GPU (qgpu_table_append_csv_file).  python scripts/csv_probe.py [sf]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyarrow.csv as pacsv
from qurious_b200 import _lib, tpch
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
cols = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]
raw = tpch.gen_lineitem(sf, device="cuda", columns=cols)
for d in (raw.cols, raw.codes):
    for c in list(d):
        d[c] = d[c].cpu()
batch = tpch.to_arrow(raw, None)[0]
path = os.path.join(tempfile.mkdtemp(prefix="qgpu_csv_"), "lineitem.tbl")
pacsv.write_csv(batch, path, pacsv.WriteOptions(include_header=False, delimiter="|", quoting_style="none"))
size = os.path.getsize(path)
ctx = _lib.Context(0)
for i in range(3):
    t0 = time.perf_counter()
    dev = _lib.DeviceTable.create(ctx, batch.schema)
    n = dev.append_csv(path, has_header=False, delimiter="|")
    dt = time.perf_counter() - t0
    dev.free()
    print(f"run {i}: {n} rows, {size / 1e6:.1f} MB of text in {dt * 1e3:.1f} ms = {size / dt / 1e9:.2f} GB/s, {n / dt / 1e6:.1f} M rows/s", flush=True)
os.remove(path)
