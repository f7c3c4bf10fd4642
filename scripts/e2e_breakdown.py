"""Where does an end-to-end step (host Arrow -> HBM -> operators -> host) spend its time?  Prints host wall
time per phase and the per-kernel CUDA-event report.  Usage: python scripts/e2e_breakdown.py [q1|q6] [sf]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from qurious_b200 import _lib, tpch  # noqa: E402
from qurious_b200.physical.plan import MemoryTable  # noqa: E402

q = sys.argv[1] if len(sys.argv) > 1 else "q1"
sf = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
ctx = _lib.Context(0)
raw = bench.gen_raw(q, sf, "cuda", 0, 1)
host = {}
for k, v in raw.items():
    for d in (v.cols, v.codes):
        for c in list(d):
            d[c] = d[c].cpu()
    host[k] = tpch.to_arrow(v, None)
del raw
torch.cuda.empty_cache()
regs = []
for k in host:
    regs += bench.pin_batches(host[k])
print("h2d bytes", sum(bench.batches_nbytes(b) for b in host.values()), "pinned buffers", len(regs), "pin failures", bench.PIN_FAILURES)
for it in range(3):
    ctx.profile(True)
    ctx.profile_report()
    t0 = time.perf_counter()
    tabs = {k: MemoryTable.try_new(b[0].schema, b) for k, b in host.items()}
    for t in tabs.values():
        t.device_table(ctx).column_bytes(0)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    plan = bench.build_plan(q, tabs)
    out = plan.execute(ctx)
    t2 = time.perf_counter()
    rep = ctx.profile_report()
    ctx.profile(False)
    print(f"iter {it}: upload {1e3 * (t1 - t0):.1f} ms, execute {1e3 * (t2 - t1):.1f} ms [{plan.last_strategy()}]")
    for r in sorted(rep, key=lambda r: -r[2])[:10]:
        print("   ", r)
    for t in tabs.values():
        t._dev.free()
