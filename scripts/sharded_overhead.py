"""Where does a sharded Q1 step spend its time compared with the single-GPU step?  One GPU, world = 1 (the all-gather
is a device copy).  Usage: [QGPU_TRACE=1] python scripts/sharded_overhead.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from qurious_b200 import _lib, tpch  # noqa: E402
from qurious_b200 import distributed as qd  # noqa: E402

ctx = _lib.Context(0)
raw = bench.gen_raw("q1", 10.0, "cuda", 0, 1)
tabs = {k: tpch.to_device_table(ctx, v) for k, v in raw.items()}
plan = bench.build_plan("q1", tabs)
sh = qd.ShardedAggregate(ctx, plan, 0, 1, all_gather=lambda o, i: o.copy_(i))


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


trace = os.environ.pop("QGPU_TRACE", None)
print("single execute_device      %.3f ms" % timeit(lambda: plan.execute_device(ctx).free()))
print("sharded execute_device     %.3f ms" % timeit(lambda: sh.execute_device().free()))
print("  partial only             %.3f ms" % timeit(lambda: sh.partial()))
