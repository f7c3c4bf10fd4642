"""N-rank correctness check over REAL NCCL / NVLink (torchrun --nproc-per-node N scripts/check_multigpu.py): every
sharded protocol must return exactly the rows of the single-table plan, which rank 0 computes on its own GPU from the
whole tables:
  * Q1 / Q6: row-range shards, fused scan + epilogue with peer exchange and merge   (qgpu_plan_execute_sharded)
  * the same through the three-call protocol over the library's NCCL all-gather   (partial_state / qgpu_comm_all_gather / execute_merged)
  * Q3: Broadcast + FinalAggregate exchange operators in one native plan          (BroadcastJoinAggregate)
  * high-cardinality group-by: peer-to-peer radix exchange                        (ExchangeGroupBy)
Exit code 0 and a final line "MULTIGPU CHECK OK" on rank 0 when everything matches."""
import os
import sys

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
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = _lib.Context(lr)
qd.init_comm(ctx)
assert ctx.comm_world() == (rank, world)
SF = float(os.environ.get("CHECK_SF", "0.5"))
ok = True


def report(name, good, extra=""):
    global ok
    flag = torch.tensor([1 if good else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"[check x{world}] {name}: {'OK' if flag.item() else 'MISMATCH'} {extra}", flush=True)
    ok = ok and bool(flag.item())


def single(query, sf):
    whole = {k: tpch.to_device_table(ctx, v) for k, v in bench.gen_raw(query, sf, "cuda", 0, 1).items()}
    return bench.build_plan(query, whole).execute(ctx)


# ---- Q1 / Q6: fused peer exchange, sync + async, and the three-call protocol over the library's NCCL all-gather ----------
for q in ("q1", "q6"):
    raw = bench.gen_raw(q, SF, "cuda", rank, world)
    tabs = {k: tpch.to_device_table(ctx, v) for k, v in raw.items()}
    plan = bench.build_plan(q, tabs)
    lo = bench.shard_range(tpch.n_lineitems(SF), rank, world)[0]
    ref = rows_of(single(q, SF)) if rank == 0 else None
    box = [ref]
    dist.broadcast_object_list(box, src=0)
    ref = box[0]
    sh = qd.ShardedAggregate(ctx, plan, lo, world)
    got = rows_of(sh.execute())
    report(f"{q} fused sharded execute", got == ref, plan.last_strategy()[-60:])
    ts = [sh.execute_device(wait=False) for _ in range(3)]       # several steps in flight
    good = True
    for t in ts:
        t.wait()
        good = good and rows_of([t.to_batch()]) == ref
        t.free()
    report(f"{q} fused sharded async x3", good)

    def nccl_all_gather(out, inp):                               # the library's communicator, on the library's stream
        ctx.comm_all_gather(inp.data_ptr(), out.data_ptr(), inp.numel())
    sh3 = qd.ShardedAggregate(ctx, plan, lo, world, all_gather=nccl_all_gather)
    report(f"{q} three-call protocol over qgpu_comm_all_gather", rows_of(sh3.execute()) == ref)

# ---- Q3: broadcast build + gather-merge ---------------------------------------------------------------------------------------
raw = bench.gen_raw("q3", SF, "cuda", rank, world)
tabs = {k: tpch.to_device_table(ctx, v) for k, v in raw.items()}
for variant in ("full broadcast", "key-range pruned broadcast"):
  prune = None if variant.startswith("full") else (0, tabs["lineitem"], tabs["lineitem"].schema.get_field_index("l_orderkey"))
  bj = qd.BroadcastJoinAggregate(ctx, tpch.q3_build_plan(tpch.Database(0.0, tabs["customer"], tabs["orders"], None)),
                                 lambda b: tpch.q3_probe_plan(b, tabs["lineitem"]), world, prune=prune)
  for rnd in range(3):                      # a learning run, then replays of the learned counts
      mine = rows_of(bj.execute())
      parts = [None] * world
      dist.all_gather_object(parts, mine)    # the result stays sharded: every group on exactly one rank
      got = sorted(r for part in parts for r in part)
      if rnd == 0:
          ref = sorted(rows_of(single("q3", SF))) if rank == 0 else None
          box = [ref]
          dist.broadcast_object_list(box, src=0)
      keys = [r[0] for r in got]
      report(f"q3 {variant} + final aggregate (native plan, run {rnd})", got == box[0] and len(keys) == len(set(keys)),
             f"{len(got)} groups, {len(mine)} here; " + bj.last_strategy[-150:])

# ---- group-by exchange (configs[3] shape at 1/250 scale) ----------------------------------------------------------------------
os.environ["QGPU_RADIX"] = "force"
rows_n, groups_n = 4_000_000, 400_000
raw = bench.gen_raw("groupby", (rows_n, groups_n), "cuda", rank, world)
tab = {k: tpch.to_device_table(ctx, v) for k, v in raw.items()}
plan = bench.build_plan("groupby", tab)
xg = qd.ExchangeGroupBy(ctx, plan, world, rank)
mine = rows_of(xg.execute())
allrows = [None] * world
dist.all_gather_object(allrows, mine)
got = sorted(r for part in allrows for r in part)
if rank == 0:
    ref = sorted(rows_of(single("groupby", (rows_n, groups_n))))
    good = len(got) == len(ref) and all(a[:5] == b[:5] and abs(a[5] - b[5]) <= 1e-12 * abs(b[5]) for a, b in zip(got, ref))
else:
    good = True
report("group-by p2p exchange", good, f"path={xg.last_path} {len(got)} groups")

dist.barrier()
if rank == 0:
    print("MULTIGPU CHECK OK" if ok else "MULTIGPU CHECK FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
