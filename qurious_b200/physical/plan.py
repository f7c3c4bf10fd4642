"""Physical operators: host-side mirror of qurious/src/physical/plan/* for the hot path.

Each class has the same name, constructor argument list and error behaviour as the reference
struct it replaces, so a test written against qurious's operators reads the same here:

  MemoryTable            qurious/src/datasource/memory.rs:20-45   (GpuMemoryTable: columns live in HBM)
  Scan                   qurious/src/physical/plan/scan.rs:12-47
  Filter                 qurious/src/physical/plan/filter.rs:12-48
  Projection             qurious/src/physical/plan/projection.rs:10-50
  NoGroupingAggregate    qurious/src/physical/plan/aggregate/no_grouping.rs:9-66
  HashAggregate          qurious/src/physical/plan/aggregate/hash.rs:110-175
  HashJoinExec, JoinFilter  qurious/src/physical/plan/join/hash_join.rs:110-385, nest_loop_join.rs:29-40

`execute()` runs the whole subtree on the GPU through the C ABI (include/qgpu.h) -- intermediate
results stay in HBM as index vectors over the base tables -- and returns `List[pyarrow.RecordBatch]`
(the reference returns `Vec<RecordBatch>`).  No compute happens in Python.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import pyarrow as pa

from .. import _lib
from ..datatypes import JoinSide, JoinType, type_triple
from .expr import AggregateExpr, PhysicalExpr

FIELD_QUALIFIERS_META_KEY = b"qurious.field_qualifiers"  # common/table_schema.rs:18


class MemoryTable:
    """`MemoryTable::try_new(schema, data)`.  Batches are uploaded to HBM on first use and stay
    resident (this is where Q1/Q6/Q3's pushed-down WHERE runs: memory.rs:69-98).

    upload_columns: optional subset of column indices to copy to the GPU (the planner knows which
    columns a query references; the reference always carries all of them)."""

    def __init__(self, schema: pa.Schema, data: Sequence[pa.RecordBatch],
                 upload_columns: Optional[Sequence[int]] = None, ctx: Optional[_lib.Context] = None):
        self.schema = schema
        self.data = list(data)
        self.upload_columns = None if upload_columns is None else list(upload_columns)
        self._ctx = ctx
        self._dev: Optional[_lib.DeviceTable] = None

    @staticmethod
    def try_new(schema: pa.Schema, data: Sequence[pa.RecordBatch], **kw) -> "MemoryTable":
        return MemoryTable(schema, data, **kw)

    @staticmethod
    def from_device_table(dev: "_lib.DeviceTable") -> "MemoryTable":
        t = MemoryTable(dev.schema, [], ctx=dev.ctx)
        t._dev = dev
        return t

    def device_table(self, ctx: _lib.Context) -> "_lib.DeviceTable":
        if self._dev is None:
            dev = _lib.DeviceTable.create(ctx, self.schema)
            for b in self.data:      # (Schema.equals first: identical schemas compare by pointer, 0.1 us instead of 10 us per batch)
                if not b.schema.equals(self.schema) and b.schema.types != self.schema.types:
                    raise _lib.QuriousError(2, "ArrowError: batch schema does not match the table schema")
            if len(self.data) > 1:
                dev.append_batches(self.data, self.upload_columns)     # one FFI call for all batches
            else:
                for b in self.data:
                    dev.append(b, self.upload_columns)
            self._dev = dev
        return self._dev

    def insert(self, batches: Sequence[pa.RecordBatch]) -> int:
        """MemoryTable::insert (memory.rs:104-111)."""
        n = 0
        for b in batches:
            self.data.append(b)
            if self._dev is not None:
                self._dev.append(b, self.upload_columns)
            n += b.num_rows
        return n


class PhysicalPlan:
    """`trait PhysicalPlan { schema, execute, children }` (physical/plan/mod.rs:25-29)."""

    schema: pa.Schema
    _last_stats: Tuple[float, int] = (0.0, 0)
    _last_strategy: str = ""

    def children(self) -> Optional[List["PhysicalPlan"]]:
        raise NotImplementedError

    def _build(self, ctx: _lib.Context, keep: list) -> ctypes.c_void_p:
        raise NotImplementedError

    def _native(self, ctx: Optional[_lib.Context] = None):
        ctx = ctx or _lib.default_context()
        keep: list = []
        return ctx, self._build(ctx, keep), keep

    @staticmethod
    def _free(ctx, keep):
        for kind, h in reversed(keep):
            if kind == "plan":
                ctx.lib.qgpu_plan_free(h)
            elif kind == "expr":
                ctx.lib.qgpu_expr_free(h)

    def _native_cached(self, ctx: Optional[_lib.Context] = None):
        """The native plan (qgpu_plan handles) is built once per (plan object, context) and re-used by later
        execute() calls: like the reference's operators, a plan is immutable and re-executable
        (physical/plan/mod.rs:25-29 takes &self)."""
        ctx = ctx or _lib.default_context()
        cached = self.__dict__.get("_native_handles")
        if cached is not None and cached[0] is ctx and ctx.handle:
            return cached
        if cached is not None:
            self.release()
        ctx, h, keep = self._native(ctx)
        self.__dict__["_native_handles"] = (ctx, h, keep)
        return ctx, h, keep

    def release(self):
        """Free the cached native plan handles."""
        cached = self.__dict__.pop("_native_handles", None)
        if cached is not None and cached[0].handle:
            self._free(cached[0], cached[2])

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def execute(self, ctx: Optional[_lib.Context] = None) -> List[pa.RecordBatch]:
        ctx, h, keep = self._native_cached(ctx)
        stream = _lib.new_stream()
        ctx.check(ctx.lib.qgpu_plan_execute(h, _lib.addr(stream)))
        self._record_stats(ctx, h)
        return _lib.read_stream(ctx, stream)

    def execute_device(self, ctx: Optional[_lib.Context] = None) -> "_lib.DeviceTable":
        """Like execute() but the result stays in HBM."""
        ctx, h, keep = self._native_cached(ctx)
        out = ctypes.c_void_p()
        nb = ctypes.c_int64()
        ctx.check(ctx.lib.qgpu_plan_execute_device(h, ctypes.byref(out), ctypes.byref(nb)))
        self._record_stats(ctx, h)
        t = _lib.DeviceTable(ctx, out, self.schema)
        t.reference_num_batches = nb.value
        return t

    def execute_device_async(self, ctx: Optional[_lib.Context] = None) -> "_lib.DeviceTable":
        """qgpu_plan_execute_device_async: returns once the kernels are queued; `DeviceTable.wait()` / `.num_rows` wait
        for the result metadata (and raise the producing kernels' errors)."""
        ctx, h, keep = self._native_cached(ctx)
        out = ctypes.c_void_p()
        ctx.check(ctx.lib.qgpu_plan_execute_device_async(h, ctypes.byref(out)))
        return _lib.DeviceTable(ctx, out, self.schema)

    def set_order_free(self, on: bool = True, ctx: Optional[_lib.Context] = None):
        """The consumer does not depend on this plan's output row order (qgpu_plan_set_order_free): lets an Inner join
        below Projection/Filter operators run as one unordered probe-scan kernel.  Not part of the reference's API."""
        ctx, h, _ = self._native_cached(ctx)
        ctx.check(ctx.lib.qgpu_plan_set_order_free(h, 1 if on else 0))
        return self

    def _record_stats(self, ctx, h):
        ms = ctypes.c_double()
        ln = ctypes.c_int64()
        ctx.lib.qgpu_plan_last_stats(h, ctypes.byref(ms), ctypes.byref(ln))
        self._last_stats = (ms.value, ln.value)
        self._last_strategy = ctx.lib.qgpu_plan_strategy(h).decode()

    def last_stats(self) -> Tuple[float, int]:
        """(device milliseconds measured with CUDA events, kernel launches) of the last execute."""
        return self._last_stats

    def last_strategy(self) -> str:
        return self._last_strategy


def _exprs_array(ctx, exprs: Sequence[PhysicalExpr], keep: list):
    hs = []
    for e in exprs:
        h = ctx.parse_expr(e)
        keep.append(("expr", h))
        hs.append(h)
    arr = (ctypes.c_void_p * max(len(hs), 1))(*[h.value for h in hs])
    return arr, hs


class Scan(PhysicalPlan):
    """`Scan::new(schema, datasource, projections, filter)` (scan.rs:22-35)."""

    def __init__(self, schema: pa.Schema, datasource: MemoryTable, projections: Optional[List[str]],
                 filter: Optional[PhysicalExpr]):
        self.schema = schema
        self.datasource = datasource
        self.projections = projections
        self.filter = filter

    def children(self):
        return None  # scan.rs:44-46

    def _build(self, ctx, keep):
        dev = self.datasource.device_table(ctx)
        fh = None
        if self.filter is not None:
            fh = ctx.parse_expr(self.filter)
            keep.append(("expr", fh))
        proj = None
        n_proj = 0
        if self.projections is not None:
            idx = []
            for name in self.projections:
                i = self.datasource.schema.get_field_index(name)
                if i < 0:
                    raise _lib.QuriousError(2, f'ArrowError: Schema error: Unable to get field named "{name}"')
                idx.append(i)
            proj = (ctypes.c_int32 * max(len(idx), 1))(*idx)
            n_proj = len(idx)
        h = ctypes.c_void_p()
        ctx.check(ctx.lib.qgpu_plan_scan(ctx.handle, dev.handle, proj, n_proj, fh, ctypes.byref(h)))
        keep.append(("plan", h))
        return h


class Filter(PhysicalPlan):
    """`Filter::new(input, predicate)` (filter.rs:18-20)."""

    def __init__(self, input: PhysicalPlan, predicate: PhysicalExpr):
        self.input = input
        self.predicate = predicate
        self.schema = input.schema

    def children(self):
        return [self.input]

    def _build(self, ctx, keep):
        ih = self.input._build(ctx, keep)
        ph = ctx.parse_expr(self.predicate)
        keep.append(("expr", ph))
        h = ctypes.c_void_p()
        ctx.check(ctx.lib.qgpu_plan_filter(ctx.handle, ih, ph, ctypes.byref(h)))
        keep.append(("plan", h))
        return h


class Projection(PhysicalPlan):
    """`Projection::new(schema, input, exprs)` (projection.rs:17-19)."""

    def __init__(self, schema: pa.Schema, input: PhysicalPlan, exprs: Sequence[PhysicalExpr]):
        self.schema = schema
        self.input = input
        self.exprs = list(exprs)

    def children(self):
        return [self.input]

    def _build(self, ctx, keep):
        ih = self.input._build(ctx, keep)
        arr, hs = _exprs_array(ctx, self.exprs, keep)
        cs = ctx.export_schema(self.schema)
        h = ctypes.c_void_p()
        try:
            ctx.check(ctx.lib.qgpu_plan_projection(ctx.handle, _lib.addr(cs), ih, arr, len(hs), ctypes.byref(h)))
        finally:
            cs.release(cs)
        keep.append(("plan", h))
        return h


class SortOptions:
    """arrow::compute::SortOptions {descending, nulls_first}."""

    def __init__(self, descending: bool = False, nulls_first: bool = True):
        self.descending, self.nulls_first = bool(descending), bool(nulls_first)


class PhyscialSortExpr:
    """`PhyscialSortExpr::new(expr, options)` (sort.rs:12-21; the reference's spelling)."""

    def __init__(self, expr: PhysicalExpr, options: SortOptions):
        self.expr, self.options = expr, options


class Sort(PhysicalPlan):
    """`Sort::new(exprs, input)` / `Sort::new_with_limit(exprs, input, limit)` (sort.rs:29-41)."""

    def __init__(self, exprs: Sequence[PhyscialSortExpr], input: PhysicalPlan, limit: Optional[int] = None):
        self.exprs = list(exprs)
        self.input = input
        self.limit = limit
        self.schema = input.schema       # sort.rs:44-46

    @staticmethod
    def new_with_limit(exprs, input, limit) -> "Sort":
        return Sort(exprs, input, limit)

    def children(self):
        return self.input.children()     # sort.rs:79-81: the reference skips the input itself

    def _build(self, ctx, keep):
        ih = self.input._build(ctx, keep)
        arr, hs = _exprs_array(ctx, [e.expr for e in self.exprs], keep)
        n = len(self.exprs)
        desc = (ctypes.c_int32 * max(n, 1))(*[1 if e.options.descending else 0 for e in self.exprs])
        nf = (ctypes.c_int32 * max(n, 1))(*[1 if e.options.nulls_first else 0 for e in self.exprs])
        h = ctypes.c_void_p()
        ctx.check(ctx.lib.qgpu_plan_sort(ctx.handle, ih, arr, desc, nf, n, -1 if self.limit is None else int(self.limit), ctypes.byref(h)))
        keep.append(("plan", h))
        return h


class Limit(PhysicalPlan):
    """`Limit::new(input, fetch, skip)` (limit.rs:15-19)."""

    def __init__(self, input: PhysicalPlan, fetch: Optional[int], skip: int):
        self.input, self.fetch, self.skip = input, fetch, int(skip)
        self.schema = input.schema       # limit.rs:23-25

    def children(self):
        return self.input.children()     # limit.rs:60-62

    def _build(self, ctx, keep):
        ih = self.input._build(ctx, keep)
        h = ctypes.c_void_p()
        ctx.check(ctx.lib.qgpu_plan_limit(ctx.handle, ih, -1 if self.fetch is None else int(self.fetch), self.skip, ctypes.byref(h)))
        keep.append(("plan", h))
        return h


def _agg_descs(ctx, aggs: Sequence[AggregateExpr], keep: list):
    descs = (_lib.qgpu_agg_desc * max(len(aggs), 1))()
    for i, a in enumerate(aggs):
        eh = ctx.parse_expr(a.expression())
        keep.append(("expr", eh))
        descs[i].op = int(a.op)
        descs[i].expr = eh.value
        t, p, s = type_triple(a.return_type)
        descs[i].return_type = _lib.qgpu_type(t, p, s)
        et = getattr(a, "expr_data_type", None)
        if et is not None:
            t, p, s = type_triple(et)
            descs[i].expr_type = _lib.qgpu_type(t, p, s)
    return descs


class _AggregateBase(PhysicalPlan):
    def _build_agg(self, ctx, keep, group_exprs, aggs):
        ih = self.input._build(ctx, keep)
        garr, ghs = _exprs_array(ctx, group_exprs, keep)
        descs = _agg_descs(ctx, aggs, keep)
        cs = ctx.export_schema(self.schema)
        h = ctypes.c_void_p()
        try:
            ctx.check(ctx.lib.qgpu_plan_aggregate(ctx.handle, _lib.addr(cs), ih, garr, len(ghs), descs, len(aggs),
                                                  ctypes.byref(h)))
        finally:
            cs.release(cs)
        keep.append(("plan", h))
        return h


class NoGroupingAggregate(_AggregateBase):
    """`NoGroupingAggregate::new(schema, input, aggr_expr)` (no_grouping.rs:16-22)."""

    def __init__(self, schema: pa.Schema, input: PhysicalPlan, aggr_expr: Sequence[AggregateExpr]):
        self.schema = schema
        self.input = input
        self.aggr_expr = list(aggr_expr)

    def children(self):
        return None  # no_grouping.rs:64-66

    def _build(self, ctx, keep):
        return self._build_agg(ctx, keep, [], self.aggr_expr)


class HashAggregate(_AggregateBase):
    """`HashAggregate::new(schema, input, group_exprs, aggregate_exprs)` (hash.rs:118-130)."""

    def __init__(self, schema: pa.Schema, input: PhysicalPlan, group_exprs: Sequence[PhysicalExpr],
                 aggregate_exprs: Sequence[AggregateExpr]):
        self.schema = schema
        self.input = input
        self.group_exprs = list(group_exprs)
        self.aggregate_exprs = list(aggregate_exprs)

    def children(self):
        return [self.input]

    def _build(self, ctx, keep):
        if not self.group_exprs:
            raise _lib.QuriousError(1, "InternalError: HashAggregate requires group expressions; "
                                       "use NoGroupingAggregate")
        return self._build_agg(ctx, keep, self.group_exprs, self.aggregate_exprs)


class JoinFilter:
    """`JoinFilter { expr, schema, column_indices }` (nest_loop_join.rs:29-40)."""

    def __init__(self, expr: PhysicalExpr, schema: pa.Schema, column_indices: Sequence[Tuple[int, JoinSide]]):
        self.expr = expr
        self.schema = schema
        self.column_indices = [(int(i), JoinSide(s)) for i, s in column_indices]


def build_join_schema(left: pa.Schema, right: pa.Schema, join_type: JoinType):
    """join/mod.rs:26-123 -> (schema, column_indices).  Host-side schema bookkeeping only; the
    library computes the same schema (qgpu_plan_schema) and tests compare the two."""
    SEP = "\x1f"
    jt = JoinType(join_type)
    if jt in (JoinType.LeftSemi, JoinType.LeftAnti):
        return (pa.schema(list(left), metadata=left.metadata),
                [(i, JoinSide.Left) for i in range(len(left))])
    ln, rn = {JoinType.Left: (False, True), JoinType.Right: (True, False),
              JoinType.Inner: (False, False), JoinType.Full: (True, True)}[jt]
    fields = [f.with_nullable(True) if ln else f for f in left] + \
             [f.with_nullable(True) if rn else f for f in right]
    idx = [(i, JoinSide.Left) for i in range(len(left))] + [(i, JoinSide.Right) for i in range(len(right))]

    def parts(s: pa.Schema):
        md = s.metadata or {}
        if FIELD_QUALIFIERS_META_KEY not in md:
            return [""] * len(s)
        p = md[FIELD_QUALIFIERS_META_KEY].decode().split(SEP)
        return p if len(p) == len(s) else [""] * len(s)

    meta = dict(left.metadata or {})
    meta[FIELD_QUALIFIERS_META_KEY] = SEP.join(parts(left) + parts(right)).encode()
    return pa.schema(fields, metadata=meta), idx


class HashJoinExec(PhysicalPlan):
    """`HashJoinExec::try_new(left, right, join_type, on, filter)` (hash_join.rs:122-146).
    The build side is always `left` (hash_join.rs:355-359)."""

    def __init__(self, left: PhysicalPlan, right: PhysicalPlan, join_type: JoinType,
                 on: Sequence[Tuple[PhysicalExpr, PhysicalExpr]], filter: Optional[JoinFilter]):
        if len(on) == 0:
            raise _lib.QuriousError(1, "InternalError: On constraints in HashJoinExec should be non-empty")
        self.left = left
        self.right = right
        self.join_type = JoinType(join_type)
        self.on = list(on)
        self.filter = filter
        self.schema, self.column_indices = build_join_schema(left.schema, right.schema, self.join_type)

    @staticmethod
    def try_new(left, right, join_type, on, filter) -> "HashJoinExec":
        return HashJoinExec(left, right, join_type, on, filter)

    def children(self):
        return [self.left, self.right]

    def _build(self, ctx, keep):
        lh = self.left._build(ctx, keep)
        rh = self.right._build(ctx, keep)
        larr, lhs = _exprs_array(ctx, [l for l, _ in self.on], keep)
        rarr, rhs = _exprs_array(ctx, [r for _, r in self.on], keep)
        fptr = None
        cs = None
        if self.filter is not None:
            fh = ctx.parse_expr(self.filter.expr)
            keep.append(("expr", fh))
            cs = ctx.export_schema(self.filter.schema)
            n = len(self.filter.column_indices)
            ci = (ctypes.c_int32 * max(n, 1))(*[i for i, _ in self.filter.column_indices])
            sd = (ctypes.c_int32 * max(n, 1))(*[int(s) for _, s in self.filter.column_indices])
            jf = _lib.qgpu_join_filter(fh.value, _lib.addr(cs), ci, sd, n)
            fptr = ctypes.byref(jf)
        h = ctypes.c_void_p()
        try:
            ctx.check(ctx.lib.qgpu_plan_hash_join(ctx.handle, lh, rh, int(self.join_type), larr, rarr, len(lhs),
                                                  fptr, ctypes.byref(h)))
        finally:
            if cs is not None:
                cs.release(cs)
        keep.append(("plan", h))
        return h


def _join_filter_struct(ctx, filter, keep):
    """-> (ctypes pointer or None, exported ArrowSchema to release or None)"""
    if filter is None:
        return None, None
    fh = ctx.parse_expr(filter.expr)
    keep.append(("expr", fh))
    cs = ctx.export_schema(filter.schema)
    n = len(filter.column_indices)
    ci = (ctypes.c_int32 * max(n, 1))(*[i for i, _ in filter.column_indices])
    sd = (ctypes.c_int32 * max(n, 1))(*[int(s) for _, s in filter.column_indices])
    jf = _lib.qgpu_join_filter(fh.value, _lib.addr(cs), ci, sd, n)
    keep.append(("raw", (jf, ci, sd)))
    return ctypes.byref(jf), cs


class NestedLoopJoinExec(PhysicalPlan):
    """`NestedLoopJoinExec::try_new(left, right, join_type, filter)` (nest_loop_join.rs:52-76): the planner's operator
    for joins without equi-conditions (planner/mod.rs:316-320).  Matched pairs are ordered by (right row, left row);
    Left / Right / Full add one batch of unmatched left rows followed by unmatched right rows (:168-226)."""

    def __init__(self, left: PhysicalPlan, right: PhysicalPlan, join_type: JoinType, filter: Optional[JoinFilter]):
        self.left = left
        self.right = right
        self.join_type = JoinType(join_type)
        self.filter = filter
        self.schema, self.column_indices = build_join_schema(left.schema, right.schema, self.join_type)

    @staticmethod
    def try_new(left, right, join_type, filter) -> "NestedLoopJoinExec":
        return NestedLoopJoinExec(left, right, join_type, filter)

    def children(self):
        return [self.left, self.right]

    def _build(self, ctx, keep):
        lh = self.left._build(ctx, keep)
        rh = self.right._build(ctx, keep)
        fptr, cs = _join_filter_struct(ctx, self.filter, keep)
        h = ctypes.c_void_p()
        try:
            ctx.check(ctx.lib.qgpu_plan_nested_loop_join(ctx.handle, lh, rh, int(self.join_type), fptr, ctypes.byref(h)))
        finally:
            if cs is not None:
                cs.release(cs)
        keep.append(("plan", h))
        return h


class CrossJoin(PhysicalPlan):
    """`CrossJoin::new(left, right)` (join/cross_join.rs:62-116): the cartesian product; what `FROM a, b` stays when no
    equi-condition links the two sides (optimizer/rule/eliminate_cross_join.rs).  Schema = left fields ++ right fields with
    the merged field qualifiers.  The library returns the rows left-row major; the reference emits one batch per (left batch,
    right batch, left row), i.e. the same rows in an order that depends on its inputs' batch boundaries (cross_join.rs:118-168)."""

    def __init__(self, left: PhysicalPlan, right: PhysicalPlan):
        self.left = left
        self.right = right
        self.schema, self.column_indices = build_join_schema(left.schema, right.schema, JoinType.Inner)

    @staticmethod
    def new(left, right) -> "CrossJoin":
        return CrossJoin(left, right)

    def children(self):
        return [self.left, self.right]

    def _build(self, ctx, keep):
        lh = self.left._build(ctx, keep)
        rh = self.right._build(ctx, keep)
        h = ctypes.c_void_p()
        ctx.check(ctx.lib.qgpu_plan_cross_join(ctx.handle, lh, rh, ctypes.byref(h)))
        keep.append(("plan", h))
        return h


class Broadcast(PhysicalPlan):
    """Exchange operator of a distributed plan (no reference counterpart: qurious is single-process): every rank executes
    `input` over its shard and receives the rows of ALL ranks, in rank order -- the build side of a broadcast join (SURVEY 8e
    "Q3 joins").  order_free: the consumer does not depend on the input's row order (it feeds a hash table)."""

    def __init__(self, input: PhysicalPlan, order_free: bool = False, prune: Optional[Tuple[int, "MemoryTable", int]] = None):
        """prune = (key column of `input`, this rank's probe-side MemoryTable, its key column): the rows feed the build side of
        an Inner equi-join against that table, so a row only has to reach the ranks whose probe-side key RANGE contains its key
        (dynamic partition pruning; implies order_free)."""
        self.input, self.order_free, self.schema, self.prune = input, bool(order_free) or prune is not None, input.schema, prune

    def children(self):
        return [self.input]

    def _build(self, ctx, keep):
        ch = self.input._build(ctx, keep)
        h = ctypes.c_void_p()
        if self.prune is not None:
            key_col, table, probe_col = self.prune
            dev = table.device_table(ctx)
            ctx.check(ctx.lib.qgpu_plan_broadcast_pruned(ctx.handle, ch, int(key_col), dev.handle, int(probe_col), ctypes.byref(h)))
            keep.append(("plan", h))
            return h
        ctx.check(ctx.lib.qgpu_plan_broadcast(ctx.handle, ch, 1 if self.order_free else 0, ctypes.byref(h)))
        keep.append(("plan", h))
        return h


class FinalAggregate(PhysicalPlan):
    """Exchange operator of a distributed plan: `input` yields PARTIAL groups per rank (groups may straddle shards); the rows
    are hash-partitioned on the first key column, exchanged all-to-all and re-aggregated with one merge operator per value
    column ("sum" for partial sums and counts, "min", "max").  The result stays sharded: every final group is returned by
    exactly one rank.  keys / values: column indices of `input`'s schema."""

    OPS = {"sum": 0, "min": 1, "max": 2}

    def __init__(self, input: PhysicalPlan, keys: Sequence[int], values: Sequence[Tuple[int, str]]):
        self.input, self.keys, self.values, self.schema = input, list(keys), list(values), input.schema

    def children(self):
        return [self.input]

    def _build(self, ctx, keep):
        ch = self.input._build(ctx, keep)
        k = (ctypes.c_int32 * len(self.keys))(*self.keys)
        n = len(self.values)
        v = (ctypes.c_int32 * max(n, 1))(*[c for c, _ in self.values])
        o = (ctypes.c_int32 * max(n, 1))(*[self.OPS[m] for _, m in self.values])
        h = ctypes.c_void_p()
        ctx.check(ctx.lib.qgpu_plan_final_aggregate(ctx.handle, ch, k, len(self.keys), v, o, n, ctypes.byref(h)))
        keep.append(("plan", h))
        return h
