"""Multi-GPU execution of aggregates over row-range shards: one process per GPU, NCCL through
torch.distributed for the plumbing.  The reference is single-process, so nothing here mirrors a reference file;
the merge itself is exact and happens on the GPU inside libqgpu (csrc/shard.cu):

    every rank:  shard-local aggregate of (Projection* <- Aggregate <- Scan(shard)), then ONE kernel: state block ->
                 peer stores into every rank's symmetric buffer (NVLink) -> epoch-flag barrier -> exact merge ->
                 finalise (AVG division, order...)                                qgpu_plan_execute_sharded
    (hosts with their own collectives: qgpu_plan_partial_state / all-gather / qgpu_plan_execute_merged)

No compute happens in Python.  `row_offset` is the global index of the shard's first row: it keeps the
first-occurrence output order identical to a single-GPU run over the whole table.
"""
from __future__ import annotations

import ctypes
import sys
from typing import Callable, List, Optional

import pyarrow as pa
import torch

from . import _lib


def shard_range(total_rows: int, rank: int, world: int):
    """Contiguous row range [lo, hi) of `rank` (rank r gets rows [r*N/G, (r+1)*N/G))."""
    return (total_rows * rank) // world, (total_rows * (rank + 1)) // world


def init_comm(ctx: _lib.Context) -> None:
    """Create the library-owned communicator of `ctx` (qgpu_comm_init): rank 0's NCCL unique id travels over the
    already-initialised torch.distributed process group (any side channel would do: a file, MPI, the host's own RPC)."""
    import torch.distributed as dist
    if ctx.comm_world()[1] > 1:
        return
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world)


class ShardedAggregate:
    """plan: (Projection|Filter)* <- HashAggregate/NoGroupingAggregate <- ... over THIS rank's shard.

    Default (all_gather=None): the whole step runs below the C ABI (qgpu_plan_execute_sharded*): shard-local aggregate,
    then ONE kernel that stores the state block into every peer's buffer over NVLink, waits on the peers' flags, merges
    and finalises.  With an explicit `all_gather` callable the three-call protocol (partial_state / caller's all-gather /
    execute_merged) is used instead -- hosts with their own collective layer, and the single-GPU emulation in the tests."""

    def __init__(self, ctx: _lib.Context, plan, row_offset: int, world: int, max_groups: int = 64,
                 all_gather: Optional[Callable[[torch.Tensor, torch.Tensor], None]] = None):
        self.ctx, self.plan, self.row_offset, self.world, self.max_groups = ctx, plan, int(row_offset), int(world), int(max_groups)
        _, self.h, _ = plan._native_cached(ctx)
        self.fused = all_gather is None
        if self.fused:
            if ctx.comm_world()[1] != self.world:
                init_comm(ctx)
            return
        n = ctypes.c_int64()
        ctx.check(ctx.lib.qgpu_plan_state_bytes(self.h, self.max_groups, ctypes.byref(n)))
        self.state_bytes = n.value
        dev = torch.device("cuda", ctx.device)
        self.state = torch.zeros(self.state_bytes, dtype=torch.uint8, device=dev)
        self.gathered = torch.zeros(self.state_bytes * self.world, dtype=torch.uint8, device=dev)
        self.stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)
        self.all_gather = all_gather
        torch.cuda.synchronize(dev)  # the zero-fills above ran on torch's stream

    def partial(self) -> torch.Tensor:
        """Run the shard-local part; returns the device state block (valid until the next call)."""
        self.ctx.check(self.ctx.lib.qgpu_plan_partial_state(self.h, self.row_offset, self.max_groups,
                                                            self.state.data_ptr(), self.state_bytes))
        return self.state

    def merge(self, gathered: torch.Tensor, n_states: int) -> List[pa.RecordBatch]:
        out = _lib.new_stream()
        self.ctx.check(self.ctx.lib.qgpu_plan_execute_merged(self.h, gathered.data_ptr(), n_states, self.max_groups,
                                                             _lib.addr(out)))
        self.plan._record_stats(self.ctx, self.h)
        return _lib.read_stream(self.ctx, out)

    def execute(self) -> List[pa.RecordBatch]:
        if self.fused:
            out = _lib.new_stream()
            self.ctx.check(self.ctx.lib.qgpu_plan_execute_sharded(self.h, self.row_offset, self.max_groups, _lib.addr(out)))
            self.plan._record_stats(self.ctx, self.h)
            return _lib.read_stream(self.ctx, out)
        self.partial()
        with torch.cuda.stream(self.stream):       # the collective is ordered after the library's stream work
            self.all_gather(self.gathered, self.state)
        return self.merge(self.gathered, self.world)

    def execute_device(self, wait: bool = True) -> _lib.DeviceTable:
        """Like execute(), the merged result stays in HBM.  wait=False (fused protocol only): returns once the kernels are
        queued; DeviceTable.wait() / .num_rows wait for the metadata."""
        out = ctypes.c_void_p()
        if self.fused:
            self.ctx.check(self.ctx.lib.qgpu_plan_execute_sharded_device(self.h, self.row_offset, self.max_groups,
                                                                         0 if wait else 1, ctypes.byref(out)))
            if wait:
                self.plan._record_stats(self.ctx, self.h)
            return _lib.DeviceTable(self.ctx, out, self.plan.schema)
        self.partial()
        with torch.cuda.stream(self.stream):
            self.all_gather(self.gathered, self.state)
        self.ctx.check(self.ctx.lib.qgpu_plan_execute_merged_device(self.h, self.gathered.data_ptr(), self.world, self.max_groups,
                                                                    ctypes.byref(out)))
        self.plan._record_stats(self.ctx, self.h)
        return _lib.DeviceTable(self.ctx, out, self.plan.schema)


# ------------------------------------------------------------------------------------------------
# hash repartition (high-cardinality group-by / join inputs) over NCCL all-to-all
# ------------------------------------------------------------------------------------------------
def partition_ids_host(keys, n_parts: int):
    """numpy mirror of csrc/shard.cu part_mix(): partition id = fmix64(key) % n_parts (tests / documentation only)."""
    import numpy as np
    x = np.asarray(keys).astype(np.int64).view(np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
    return (x % np.uint64(n_parts)).astype(np.int64)


class _DevBuf:
    """Zero-copy view of a library-owned device buffer as a torch byte tensor (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def column_bytes_tensor(table: _lib.DeviceTable, col: int):
    """-> (uint8 tensor aliasing the column's Arrow-layout value buffer in HBM, value width in bytes)."""
    ptr, nbytes, width = table.column_device_buffer(col)
    dev = torch.device("cuda", table.ctx.device)
    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=dev), width
    return torch.as_tensor(_DevBuf(ptr, nbytes), device=dev), width


def table_from_tensors(ctx: _lib.Context, schema: pa.Schema, tensors: List[torch.Tensor], n_rows: int) -> _lib.DeviceTable:
    """HBM-resident table from one device tensor per (fixed-width, NULL-free) column; the library copies the buffers
    (qgpu_table_append_device), so the tensors may be reused afterwards."""
    from pyarrow.cffi import ffi
    keep, children = [], []
    for t in tensors:
        cb = ffi.new("const void*[]", [ffi.NULL, ffi.cast("void*", t.data_ptr() if t.numel() else 0)])
        ca = ffi.new("struct ArrowArray*")
        ca.length, ca.null_count, ca.offset, ca.n_buffers, ca.n_children = n_rows, 0, 0, 2, 0
        ca.buffers, ca.release = cb, ffi.NULL
        keep += [cb, ca]
        children.append(ca)
    top = ffi.new("struct ArrowArray*")
    kids = ffi.new("struct ArrowArray*[]", children)
    topbufs = ffi.new("const void*[]", [ffi.NULL])
    top.length, top.null_count, top.offset, top.n_buffers, top.n_children = n_rows, 0, 0, 1, len(children)
    top.buffers, top.children, top.release = topbufs, kids, ffi.NULL
    dev = _lib.DeviceTable.create(ctx, schema)
    if n_rows > 0:
        dev.append_device_struct(_lib.addr(top))
        dev.column_bytes(0)  # consolidate + synchronise the library's D2D copies before the tensors go away
    return dev


def exchange_columns(columns: List[torch.Tensor], widths: List[int], send_rows: List[int], all_to_all: Callable,
                     device=None):
    """The exchange step of a hash repartition, device-agnostic (NCCL on the GPU, gloo in the CPU tests).
    columns[c]: uint8 tensor holding this rank's rows of column c grouped by destination rank; send_rows[p]: rows
    going to rank p.  -> (received byte tensors, rows received from every rank)."""
    send = torch.tensor(send_rows, dtype=torch.int64, device=device)
    recv = torch.empty_like(send)
    all_to_all(recv, send, None, None)
    recv_rows = [int(x) for x in recv.tolist()]
    n_recv = sum(recv_rows)
    outs = []
    for col, w in zip(columns, widths):
        out = torch.empty(n_recv * w, dtype=torch.uint8, device=col.device)
        all_to_all(out, col, [r * w for r in recv_rows], [r * w for r in send_rows])
        outs.append(out)
    return outs, recv_rows


def _dist_all_to_all(out, inp, out_splits, in_splits):
    import torch.distributed as dist
    dist.all_to_all_single(out, inp, out_splits, in_splits)


def hash_repartition(ctx: _lib.Context, table: _lib.DeviceTable, key_col: int, world: int,
                     all_to_all: Optional[Callable] = None) -> _lib.DeviceTable:
    """Every rank: partition its rows by fmix64(key) % world on the GPU (csrc/shard.cu: k_part_hist/k_part_scatter),
    exchange the per-destination slices of every column with NCCL all-to-all (ncclSend/ncclRecv over NVLink), return
    the received rows as a new HBM-resident table.  Afterwards equal keys live on exactly one rank, so a purely
    local aggregate is exact and the whole result is the concatenation of the ranks' results."""
    import os
    import time
    trace = bool(os.environ.get("QGPU_TRACE"))
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)
    t0 = time.perf_counter()

    def mark(what):
        nonlocal t0
        if trace:
            stream.synchronize()
            t1 = time.perf_counter()
            print(f"[qgpu trace] repartition: {what:<28s} {1e3 * (t1 - t0):9.3f} ms", file=sys.stderr)
            t0 = t1
    parted, offs = table.hash_partition(key_col, world)
    mark("partition kernels")
    cols, widths = [], []
    for c in range(len(table.schema)):
        t, w = column_bytes_tensor(parted, c)
        cols.append(t)
        widths.append(w)
    with torch.cuda.stream(stream):                # ordered after the library's partition kernels
        outs, recv_rows = exchange_columns(cols, widths, [offs[i + 1] - offs[i] for i in range(world)],
                                           all_to_all or _dist_all_to_all, device=dev)
    stream.synchronize()
    mark("all-to-all")
    res = table_from_tensors(ctx, table.schema, outs, sum(recv_rows))
    mark("received table")
    parted.free()
    return res


class ExchangeGroupBy:
    """High-cardinality group-by over row-range shards (BASELINE.json configs[3]): ONE kernel partitions this rank's rows
    by key hash and stores every tuple straight into its owner's HBM over NVLink (csrc/radix_agg.cuh k_radix_scatter<1>
    with peer destinations) -- no send buffers, no NCCL all-to-all; NCCL only carries the 17 KB sketches, the 80-byte
    buffer handles and the barrier.  Afterwards equal keys live on one rank, which aggregates them locally; the whole
    result is the concatenation of the ranks' results.  Falls back (collectively) on hash_repartition + a local
    aggregate when the plan is not eligible or a skewed bucket overflows."""

    HANDLE_BYTES = 80

    def __init__(self, ctx: _lib.Context, plan, world: int, rank: int, collectives=None):
        self.ctx, self.plan, self.world, self.rank = ctx, plan, int(world), int(rank)
        _, self.h, _ = plan._native_cached(ctx)
        self.dev = torch.device("cuda", ctx.device)
        self.stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=self.dev)
        self.coll = collectives or _DistCollectives()
        self.last_path = ""
        self._gstats = None

    # -- the library stages -------------------------------------------------------------------------------------
    def keystats(self) -> List[int]:
        st = (ctypes.c_int64 * 8)()
        nk = ctypes.c_int32()
        self.ctx.check(self.ctx.lib.qgpu_plan_exchange_keystats(self.h, st, ctypes.byref(nk)))
        return list(st)[:2 * nk.value]

    def sketch(self, global_stats: List[int]):
        """-> (device byte tensor with this rank's sketch, eligible)"""
        st = (ctypes.c_int64 * 8)(*global_stats)
        ptr, n, ok = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int32()
        self.ctx.check(self.ctx.lib.qgpu_plan_exchange_sketch(self.h, st, ctypes.byref(ptr), ctypes.byref(n), ctypes.byref(ok)))
        if not ok.value:
            return None, False
        return torch.as_tensor(_DevBuf(ptr.value, n.value), device=self.dev), True

    def global_keystats(self) -> List[int]:
        """(min, max) of every key over all ranks; the table is immutable, so this is exchanged once."""
        if self._gstats is None:
            mine = self.keystats()
            t = torch.tensor([(-v if i % 2 == 0 else v) for i, v in enumerate(mine)], dtype=torch.int64, device=self.dev)
            t = self.coll.all_reduce_max(t)            # max of (-min, max)
            self._gstats = [(-int(v) if i % 2 == 0 else int(v)) for i, v in enumerate(t.tolist())]
        return self._gstats

    def prepare(self, gathered_host: torch.Tensor):
        handles = torch.zeros(6 * self.HANDLE_BYTES, dtype=torch.uint8)
        nh, ok = ctypes.c_int32(), ctypes.c_int32()
        self.ctx.check(self.ctx.lib.qgpu_plan_exchange_prepare(self.h, gathered_host.data_ptr(), self.world, self.rank,
                                                                handles.data_ptr(), ctypes.byref(nh), ctypes.byref(ok)))
        return handles[:nh.value * self.HANDLE_BYTES], bool(ok.value)

    def scatter(self, all_handles_host: torch.Tensor):
        self.ctx.check(self.ctx.lib.qgpu_plan_exchange_scatter(self.h, all_handles_host.data_ptr()))

    def finish(self) -> int:
        ovf = ctypes.c_int32()
        self.ctx.check(self.ctx.lib.qgpu_plan_exchange_finish(self.h, ctypes.byref(ovf)))
        return ovf.value

    # -- the collective sequence -----------------------------------------------------------------------------------------
    def execute_device(self) -> _lib.DeviceTable:
        gstats = self.global_keystats()
        with torch.cuda.stream(self.stream):        # ordered after the library's sketch kernel
            sk, ok = self.sketch(gstats)
            if ok:
                gathered = self.coll.all_gather_bytes(sk, self.world).cpu()
        if ok:
            handles, ok = self.prepare(gathered)
        if ok:
            with torch.cuda.stream(self.stream):
                all_h = self.coll.all_gather_bytes(handles.to(self.dev), self.world).cpu()
            self.scatter(all_h)
            self.coll.barrier()                      # every rank's peer stores have completed
            failed = self.finish()
            if self.coll.any_nonzero(failed, self.dev):
                if not failed:
                    self.plan.execute_device(self.ctx).free()      # drop this rank's (unused) result
                ok = False
        if ok:
            self.last_path = "exchange"
            return self.plan.execute_device(self.ctx)
        return self._fallback()

    def _fallback(self) -> _lib.DeviceTable:
        from .physical.plan import HashAggregate, MemoryTable, Projection, Scan
        self.last_path = "hash_repartition"
        node, chain = self.plan, []
        while isinstance(node, Projection):
            chain.append(node)
            node = node.input
        if not isinstance(node, HashAggregate) or not isinstance(node.input, Scan):
            raise _lib.QuriousError(1, "InternalError: exchange fallback needs (Projection)* <- HashAggregate <- Scan")
        scan = node.input
        key = node.group_exprs[0]
        table = scan.datasource.device_table(self.ctx)
        recv = hash_repartition(self.ctx, table, key.index, self.world, self.coll.all_to_all)
        p = HashAggregate(node.schema, Scan(scan.schema, MemoryTable.from_device_table(recv), scan.projections, scan.filter),
                          node.group_exprs, node.aggregate_exprs)
        for pr in reversed(chain):
            p = Projection(pr.schema, p, pr.exprs)
        out = p.execute_device(self.ctx)
        p.release()
        recv.free()
        return out

    def execute(self) -> List[pa.RecordBatch]:
        t = self.execute_device()
        out = [t.to_batch()] if t.num_rows > 0 else []
        t.free()
        return out


class _DistCollectives:
    """torch.distributed (NCCL) plumbing of ExchangeGroupBy; tests substitute an in-process emulation."""

    def all_gather_bytes(self, t: torch.Tensor, world: int) -> torch.Tensor:
        import torch.distributed as dist
        out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous())
        return out

    def barrier(self):
        import torch.distributed as dist
        dist.barrier()

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    def any_nonzero(self, v: int, dev) -> bool:
        import torch.distributed as dist
        t = torch.tensor([1 if v else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return bool(t.item())

    def all_to_all(self, out, inp, out_splits, in_splits):
        _dist_all_to_all(out, inp, out_splits, in_splits)


# ------------------------------------------------------------------------------------------------
# group-by results that straddle shards (SURVEY 8e "Q3 joins"): all-gather the per-rank result rows, re-aggregate
# ------------------------------------------------------------------------------------------------
def merge_spec_of(plan):
    """(key output columns, [(output column, merge kind)]) of `Projection(plain columns)? <- HashAggregate`.
    SUM and COUNT partials merge by SUM, MIN by MIN, MAX by MAX (exact for integers/decimals: wrapping adds are
    associative); AVG has no mergeable output (use ShardedAggregate, which carries (sum, count))."""
    from .physical.expr import (AvgAggregateExpr, Column, CountAggregateExpr, MaxAggregateExpr, MinAggregateExpr,
                                SumAggregateExpr)
    from .physical.plan import HashAggregate, Projection
    proj = None
    node = plan
    if isinstance(node, Projection):
        proj, node = node, node.input
    if not isinstance(node, HashAggregate):
        raise _lib.QuriousError(1, "InternalError: gather-merge needs (Projection <-) HashAggregate")
    n_keys = len(node.group_exprs)
    kinds = []
    for a in node.aggregate_exprs:
        if isinstance(a, (SumAggregateExpr, CountAggregateExpr)):
            kinds.append("sum")
        elif isinstance(a, MinAggregateExpr):
            kinds.append("min")
        elif isinstance(a, MaxAggregateExpr):
            kinds.append("max")
        else:
            raise _lib.QuriousError(1, f"InternalError: {type(a).__name__} partials cannot be merged from result rows")
    src = list(range(n_keys + len(kinds)))
    if proj is not None:
        if not all(isinstance(e, Column) for e in proj.exprs):
            raise _lib.QuriousError(1, "InternalError: gather-merge needs a Projection of plain columns")
        src = [e.index for e in proj.exprs]
    keys = [o for o, s in enumerate(src) if s < n_keys]
    aggs = [(o, kinds[s - n_keys]) for o, s in enumerate(src) if s >= n_keys]
    return keys, aggs


class GatherMergeAggregate:
    """Row-range shards whose groups can straddle shard boundaries and are too many for fixed-size state blocks
    (Q3: ~114k x SF groups, lineitem sorted by l_orderkey so at most world-1 groups straddle):
        every rank:  the whole local plan, result kept in HBM                    plan.execute_device
        NCCL:        all-gather of the result columns (ragged: counts first)      dist.all_gather
        every rank:  re-aggregate the gathered rows by the same keys             HashAggregate over the gathered table
    Output column order = the plan's."""

    def __init__(self, ctx: _lib.Context, plan, world: int, all_gather_ragged: Optional[Callable] = None):
        self.ctx, self.plan, self.world = ctx, plan, int(world)
        self.keys, self.aggs = merge_spec_of(plan)
        self.gather = all_gather_ragged or _dist_all_gather_ragged
        self.stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=torch.device("cuda", ctx.device))
        self._merge_plan = None
        self._merge_table = None

    def _make_merge_plan(self, table):
        from .physical.expr import Column, MaxAggregateExpr, MinAggregateExpr, SumAggregateExpr
        from .physical.plan import HashAggregate, MemoryTable, Projection, Scan
        schema = self.plan.schema
        mk = {"sum": SumAggregateExpr, "min": MinAggregateExpr, "max": MaxAggregateExpr}
        fields = [schema.field(o) for o in self.keys] + [schema.field(o) for o, _ in self.aggs]
        agg = HashAggregate(pa.schema(fields), Scan(schema, MemoryTable.from_device_table(table), None, None),
                            [Column(schema.field(o).name, o) for o in self.keys],
                            [mk[k](Column(schema.field(o).name, o), schema.field(o).type) for o, k in self.aggs])
        order = self.keys + [o for o, _ in self.aggs]          # position in agg output -> plan output column
        back = [order.index(o) for o in range(len(schema))]
        return Projection(schema, agg, [Column(schema.field(o).name, back[o]) for o in range(len(schema))])

    def execute_device(self) -> _lib.DeviceTable:
        local = self.plan.execute_device(self.ctx)
        cols, widths = [], []
        for c in range(len(self.plan.schema)):
            t, w = column_bytes_tensor(local, c)
            cols.append(t)
            widths.append(w)
        with torch.cuda.stream(self.stream):
            outs, n = self.gather(cols, widths, local.num_rows, self.world)
        self.stream.synchronize()
        gathered = table_from_tensors(self.ctx, self.plan.schema, outs, n)
        local.free()
        merged = self._make_merge_plan(gathered).execute_device(self.ctx)
        gathered.free()
        return merged

    def execute(self) -> List[pa.RecordBatch]:
        t = self.execute_device()
        out = [t.to_batch()] if t.num_rows > 0 else []
        t.free()
        return out


def all_gather_table(ctx: _lib.Context, table: _lib.DeviceTable, world: int,
                     all_gather_ragged: Optional[Callable] = None) -> _lib.DeviceTable:
    """The rows of every rank's HBM-resident table (fixed-width, NULL-free columns), in rank order, on every rank."""
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=torch.device("cuda", ctx.device))
    cols, widths = [], []
    for c in range(len(table.schema)):
        t, w = column_bytes_tensor(table, c)
        cols.append(t)
        widths.append(w)
    with torch.cuda.stream(stream):
        outs, n = (all_gather_ragged or _dist_all_gather_ragged)(cols, widths, table.num_rows, world)
    stream.synchronize()
    return table_from_tensors(ctx, table.schema, outs, n)


class BroadcastJoinAggregate:
    """Join + aggregate with BOTH inputs row-range sharded (SURVEY 8e "Q3 joins", broadcast variant).

    Default: ONE native plan per rank, executed below the C ABI (csrc/exchange.cu):
        FinalAggregate <- probe_plan_of( Broadcast(build_plan) )
    Broadcast: every rank runs the build-side sub-plan over its shard (under its own speculation scope: no host round
    trips after the first execution) and the rows of all ranks reach every GPU in one grouped NCCL exchange, integer
    statistics included; the probe plan (fused probe + aggregate kernel) then runs over this rank's shard of the fact
    table; FinalAggregate hash-partitions the partial groups (groups may straddle shard boundaries), exchanges them
    all-to-all and re-aggregates.  The result STAYS SHARDED: execute() returns this rank's final groups, every group on
    exactly one rank.  probe_plan_of(build: PhysicalPlan | MemoryTable) -> (Projection <-) HashAggregate plan.

    With an explicit `all_gather_ragged` callable the older host-driven protocol runs instead (all-gather of the build
    rows, probe, all-gather of the result rows + re-aggregation: every rank gets the whole result) -- hosts with their own
    collective layer and the single-GPU emulation in the tests."""

    def __init__(self, ctx: _lib.Context, build_plan, probe_plan_of: Callable, world: int,
                 all_gather_ragged: Optional[Callable] = None, prune=None):
        """prune = (key column of build_plan's output, this rank's probe-side MemoryTable, its key column): Broadcast's
        key-range pruning (every rank receives only the build rows its probe shard can match)."""
        self.ctx, self.build_plan, self.probe_plan_of, self.world = ctx, build_plan, probe_plan_of, int(world)
        self.gather = all_gather_ragged
        self.last_strategy = ""
        self.plan = None
        if all_gather_ragged is None:
            from .physical.plan import Broadcast, FinalAggregate
            if self.world > 1 and ctx.comm_world()[1] != self.world:
                init_comm(ctx)
            probe = probe_plan_of(Broadcast(build_plan, order_free=True, prune=prune))
            keys, aggs = merge_spec_of(probe)
            self.plan = FinalAggregate(probe, keys, aggs)

    def execute_device(self) -> _lib.DeviceTable:
        if self.plan is not None:
            out = self.plan.execute_device(self.ctx)
            self.last_strategy = self.plan.last_strategy()
            return out
        from .physical.plan import MemoryTable
        self.build_plan.set_order_free(True, self.ctx)      # the rows feed an all-gather and a hash table
        local = self.build_plan.execute_device(self.ctx)
        build = all_gather_table(self.ctx, local, self.world, self.gather)
        local.free()
        probe = self.probe_plan_of(MemoryTable.from_device_table(build))
        out = GatherMergeAggregate(self.ctx, probe, self.world, self.gather).execute_device()
        self.last_strategy = "broadcast-build[%s] -> %s" % (self.build_plan.last_strategy(), probe.last_strategy())
        probe.release()
        build.free()
        return out

    def execute(self) -> List[pa.RecordBatch]:
        t = self.execute_device()
        out = [t.to_batch()] if t.num_rows > 0 else []
        t.free()
        return out

    def release(self):
        if self.plan is not None:
            self.plan.release()


def _dist_all_gather_ragged(cols, widths, n_rows, world):
    """Ragged all-gather of a set of columns (byte tensors, `widths[c]` bytes per row): every rank receives every rank's
    rows, in rank order.  TWO collectives whatever the column count: the row counts, then ONE equal-size all-gather of
    a buffer that packs all columns of this rank, each padded to the largest shard (result rows are few); the ranks'
    slices are then compacted per column."""
    import torch.distributed as dist
    dev = cols[0].device if cols else None
    mine = torch.tensor([n_rows], dtype=torch.int64, device=dev)
    counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine)
    rows = [int(x) for x in counts.tolist()]
    mx = max(rows) if rows else 0
    offs, block = [], 0                       # byte offset of every column inside one rank's block
    for w in widths:
        offs.append(block)
        block += ((mx * w + 15) // 16) * 16
    if not cols or block == 0:
        return [torch.empty(0, dtype=torch.uint8, device=dev) for _ in cols], sum(rows)
    mine_p = torch.zeros(block, dtype=torch.uint8, device=dev)
    for t, o in zip(cols, offs):
        mine_p[o:o + t.numel()] = t
    allp = torch.empty(world * block, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allp, mine_p)
    outs = []
    for w, o in zip(widths, offs):
        outs.append(torch.cat([allp[r * block + o:r * block + o + rows[r] * w] for r in range(world)]))
    return outs, sum(rows)
