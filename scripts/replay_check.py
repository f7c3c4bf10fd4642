"""Re-execution paths of the join pipeline (learning run, speculative replay, asynchronous replay) must return the rows
of the first execution; run under compute-sanitizer to check the kernels' memory accesses:
    compute-sanitizer --tool memcheck python scripts/replay_check.py [SF]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qurious_b200 import _lib, tpch  # noqa: E402
from tests.cases import rows_of  # noqa: E402

sf = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
ctx = _lib.default_context()
db = tpch.generate(sf)
for q in ("q3", "q1", "q6"):
    plan = getattr(tpch, q + "_plan")(db)
    first = sorted(rows_of(plan.execute(ctx)))
    print(q, len(first), "rows;", plan.last_strategy()[:150], flush=True)
    for i in range(3):
        again = sorted(rows_of(plan.execute(ctx)))
        assert again == first, f"{q}: replay {i} differs"
    for i in range(3):
        t = plan.execute_device_async(ctx)
        t2 = plan.execute_device_async(ctx)          # two executions in flight
        for x in (t, t2):
            x.wait()
            got = sorted(rows_of([x.to_batch()] if x.num_rows else []))
            assert got == first, f"{q}: async replay {i} differs"
            x.free()
    print(q, "replays OK", flush=True)
print("launches", ctx.kernel_launches())
