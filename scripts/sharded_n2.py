"""Where does the sharded Q1 step go at N > 1?  torchrun --nproc-per-node N scripts/sharded_n2.py
Device-timed (CUDA events on the library stream), max over ranks, rank 0 prints."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from qurious_b200 import _lib, tpch  # noqa: E402
from qurious_b200 import distributed as qd  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = _lib.Context(lr)
stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=torch.device("cuda", lr))
raw = bench.gen_raw("q1", 10.0 * world, "cuda", rank, world)
tabs = {k: tpch.to_device_table(ctx, v) for k, v in raw.items()}
plan = bench.build_plan("q1", tabs)
lo, hi = bench.shard_range(tpch.n_lineitems(10.0 * world), rank, world)
sh = qd.ShardedAggregate(ctx, plan, lo, world)
small = torch.zeros(4096, dtype=torch.uint8, device="cuda")
small_all = torch.zeros(4096 * world, dtype=torch.uint8, device="cuda")


def timeit(name, fn, n=40):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-44s %.3f ms" % (name, t.item()), flush=True)


def ag_only():
    with torch.cuda.stream(stream):
        dist.all_gather_into_tensor(small_all, small)


def partial_ag():
    sh.partial()
    with torch.cuda.stream(stream):
        sh.all_gather(sh.gathered, sh.state)


timeit("single-GPU plan (local shard only)", lambda: plan.execute_device(ctx).free())
timeit("partial only", lambda: sh.partial())
timeit("all_gather_into_tensor only (4 KB)", ag_only)
timeit("partial + all-gather", partial_ag)
timeit("sharded execute_device (full step)", lambda: sh.execute_device().free())
dist.destroy_process_group()
