"""Times the staged ingest of Q1's lineitem columns (SF given) per phase: QGPU_INGEST_TRACE=1 python scripts/ingest_probe.py 10"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyarrow as pa
import torch
from qurious_b200 import _lib, tpch
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
cols = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]
raw = tpch.gen_lineitem(sf, device="cuda", columns=cols)
for d in (raw.cols, raw.codes):
    for c in list(d):
        d[c] = d[c].cpu()
full = tpch.to_arrow(raw, None)
ctx = _lib.Context(0)
small = [full[0].slice(o, 1024) for o in range(0, full[0].num_rows, 1024)]
for label, batches, threads, narrow in [("single", full, 0, 1), ("single", full, 0, 1), ("single t8", full, 8, 1), ("single t12", full, 12, 1), ("single nonarrow", full, 0, 0),
                                        ("1024 stream", small, 0, 1), ("1024 stream", small, 0, 1)]:
    ctx.set_option("ingest_threads", threads)
    ctx.set_option("ingest_host_narrow", narrow)
    t0 = time.perf_counter()
    dev = _lib.DeviceTable.create(ctx, full[0].schema)
    if len(batches) > 1:
        dev.append_batches(batches)
    else:
        dev.append(batches[0])
    t1 = time.perf_counter()
    dev.flush()
    t2 = time.perf_counter()
    dev.free()
    t3 = time.perf_counter()
    print(f"{label}: append {1e3*(t1-t0):.1f} ms, flush {1e3*(t2-t1):.1f} ms, free {1e3*(t3-t2):.1f} ms", flush=True)
