"""N-GPU correctness check of the sharded Q3 path over real NCCL (torchrun --nproc-per-node N scripts/check_n2_q3.py):
BroadcastJoinAggregate over row-range shards of orders / lineitem must return exactly the rows of the single-table plan
(computed by rank 0 on its own GPU from the whole tables).  Also times a few steps."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from qurious_b200 import _lib, tpch  # noqa: E402
from qurious_b200 import distributed as qd  # noqa: E402
from tests.cases import rows_of  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = _lib.Context(lr)
SF = float(os.environ.get("CHECK_SF", "1.0"))
raw = bench.gen_raw("q3", SF, "cuda", rank, world)
tabs = {k: tpch.to_device_table(ctx, v) for k, v in raw.items()}
bj = qd.BroadcastJoinAggregate(ctx, tpch.q3_build_plan(tpch.Database(0.0, tabs["customer"], tabs["orders"], None)),
                               lambda b: tpch.q3_probe_plan(b, tabs["lineitem"]), world)
got = sorted(rows_of(bj.execute()))
ok = True
if rank == 0:
    whole = {k: tpch.to_device_table(ctx, v) for k, v in bench.gen_raw("q3", SF, "cuda", 0, 1).items()}
    ref = sorted(rows_of(bench.build_plan("q3", whole).execute(ctx)))
    ok = got == ref
    print(f"sharded Q3 over NCCL, world {world}, SF {SF:g}: {len(got)} groups, equal to the single-table plan: {ok}", flush=True)
for _ in range(3):
    bj.execute_device().free()
dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    bj.execute_device().free()
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("step %.3f ms (wall, %d ranks)" % ((time.perf_counter() - t0) / 10 * 1e3, world), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
