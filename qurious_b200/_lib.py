"""ctypes binding of libqgpu.so (include/qgpu.h).

This is the Python stand-in for the Rust `extern "C"` shim described in INTEGRATION.md: it only
marshals Arrow C Data Interface structs and the expression IR across the C ABI.  There is no compute
and no fallback here: if the shared library is missing or no CUDA device is present the operators
raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

import pyarrow as pa
from pyarrow.cffi import ffi as _ffi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqgpu.so")

# every symbol include/qgpu.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "qgpu_init", "qgpu_shutdown", "qgpu_last_error", "qgpu_set_compat", "qgpu_kernel_launches",
    "qgpu_ctx_stream", "qgpu_profile_enable", "qgpu_profile_report",
    "qgpu_set_option", "qgpu_table_append_stream", "qgpu_table_flush", "qgpu_jit_compile", "qgpu_counter", "qgpu_table_append_csv", "qgpu_table_append_csv_file",
    "qgpu_table_create", "qgpu_table_append", "qgpu_table_append_device", "qgpu_table_num_rows",
    "qgpu_table_num_batches", "qgpu_table_column_bytes", "qgpu_table_schema", "qgpu_table_export",
    "qgpu_table_free", "qgpu_expr_parse", "qgpu_expr_free", "qgpu_plan_scan", "qgpu_plan_filter",
    "qgpu_plan_projection", "qgpu_plan_sort", "qgpu_plan_limit", "qgpu_plan_aggregate", "qgpu_plan_hash_join", "qgpu_plan_nested_loop_join", "qgpu_plan_cross_join", "qgpu_plan_broadcast", "qgpu_plan_broadcast_pruned", "qgpu_plan_final_aggregate", "qgpu_plan_schema",
    "qgpu_plan_execute", "qgpu_plan_execute_device", "qgpu_plan_last_stats", "qgpu_plan_strategy",
    "qgpu_plan_free", "qgpu_plan_state_bytes", "qgpu_plan_partial_state", "qgpu_plan_execute_merged", "qgpu_plan_execute_merged_device", "qgpu_plan_set_order_free",
    "qgpu_release_cached_memory", "qgpu_table_hash_partition", "qgpu_table_column_device_buffer",
    "qgpu_plan_exchange_keystats", "qgpu_plan_exchange_sketch", "qgpu_plan_exchange_prepare", "qgpu_plan_exchange_scatter", "qgpu_plan_exchange_finish",
    "qgpu_plan_execute_device_async", "qgpu_table_wait",
    "qgpu_comm_unique_id", "qgpu_comm_init", "qgpu_comm_init_local", "qgpu_comm_destroy", "qgpu_comm_world", "qgpu_comm_all_gather",
    "qgpu_comm_all_to_all", "qgpu_comm_barrier", "qgpu_plan_execute_sharded", "qgpu_plan_execute_sharded_device",
]

STATUS_KIND = {1: "InternalError", 2: "ArrowError", 3: "CudaError", 4: "NcclError", 5: "OutOfMemory"}


class QuriousError(RuntimeError):
    """Mirror of qurious::error::Error (qurious/src/error.rs:41-53)."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code
        self.kind = STATUS_KIND.get(code, "InternalError")


class qgpu_type(ctypes.Structure):
    _fields_ = [("id", ctypes.c_uint8), ("precision", ctypes.c_uint8), ("scale", ctypes.c_int8)]


class qgpu_csv_options(ctypes.Structure):
    _fields_ = [("has_header", ctypes.c_uint8), ("delimiter", ctypes.c_uint8), ("quote", ctypes.c_uint8), ("escape", ctypes.c_uint8)]


class qgpu_agg_desc(ctypes.Structure):
    _fields_ = [("op", ctypes.c_int32), ("expr", ctypes.c_void_p), ("return_type", qgpu_type),
                ("expr_type", qgpu_type)]


class qgpu_join_filter(ctypes.Structure):
    _fields_ = [("expr", ctypes.c_void_p), ("schema", ctypes.c_void_p),
                ("column_index", ctypes.POINTER(ctypes.c_int32)),
                ("column_side", ctypes.POINTER(ctypes.c_int32)), ("n_columns", ctypes.c_int32)]


_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen libqgpu.so and declare prototypes.  Raises if the CUDA extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QuriousError(3, f"libqgpu.so not found at {LIB_PATH}: build it with "
                              f"`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    P = ctypes.POINTER
    lib.qgpu_init.argtypes = [P(ctypes.c_int), ctypes.c_int, P(vp)]
    lib.qgpu_init.restype = ctypes.c_int
    lib.qgpu_shutdown.argtypes = [vp]
    lib.qgpu_shutdown.restype = None
    lib.qgpu_last_error.argtypes = [vp]
    lib.qgpu_last_error.restype = ctypes.c_char_p
    lib.qgpu_set_compat.argtypes = [vp, ctypes.c_char_p, ctypes.c_int]
    lib.qgpu_kernel_launches.argtypes = [vp]
    lib.qgpu_kernel_launches.restype = i64
    lib.qgpu_ctx_stream.argtypes = [vp]
    lib.qgpu_ctx_stream.restype = vp
    lib.qgpu_profile_enable.argtypes = [vp, ctypes.c_int]
    lib.qgpu_profile_report.argtypes = [vp, ctypes.c_char_p, i64]
    lib.qgpu_profile_report.restype = i64
    lib.qgpu_table_create.argtypes = [vp, vp, P(vp)]
    lib.qgpu_table_append.argtypes = [vp, vp, P(i32), i32]
    lib.qgpu_table_append_device.argtypes = [vp, vp]
    lib.qgpu_table_append_stream.argtypes = [vp, vp, P(i32), i32, P(i64)]
    lib.qgpu_table_flush.argtypes = [vp]
    lib.qgpu_table_append_csv.argtypes = [vp, ctypes.c_char_p, i64, P(qgpu_csv_options), P(i32), i32, P(i64)]
    lib.qgpu_table_append_csv_file.argtypes = [vp, ctypes.c_char_p, P(qgpu_csv_options), P(i32), i32, P(i64)]
    lib.qgpu_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    lib.qgpu_jit_compile.argtypes = [P(ctypes.c_uint64), ctypes.c_uint32, ctypes.c_char_p, i64]
    lib.qgpu_jit_compile.restype = i64
    lib.qgpu_counter.argtypes = [vp, ctypes.c_char_p]
    lib.qgpu_counter.restype = i64
    lib.qgpu_table_num_rows.argtypes = [vp]
    lib.qgpu_table_num_rows.restype = i64
    lib.qgpu_table_num_batches.argtypes = [vp]
    lib.qgpu_table_num_batches.restype = i64
    lib.qgpu_table_column_bytes.argtypes = [vp, i32]
    lib.qgpu_table_column_bytes.restype = i64
    lib.qgpu_table_schema.argtypes = [vp, vp]
    lib.qgpu_table_export.argtypes = [vp, vp, vp]
    lib.qgpu_table_free.argtypes = [vp]
    lib.qgpu_table_free.restype = None
    lib.qgpu_expr_parse.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t, P(vp)]
    lib.qgpu_expr_free.argtypes = [vp]
    lib.qgpu_expr_free.restype = None
    lib.qgpu_plan_scan.argtypes = [vp, vp, P(i32), i32, vp, P(vp)]
    lib.qgpu_plan_filter.argtypes = [vp, vp, vp, P(vp)]
    lib.qgpu_plan_projection.argtypes = [vp, vp, vp, P(vp), i32, P(vp)]
    lib.qgpu_plan_aggregate.argtypes = [vp, vp, vp, P(vp), i32, P(qgpu_agg_desc), i32, P(vp)]
    lib.qgpu_plan_hash_join.argtypes = [vp, vp, vp, i32, P(vp), P(vp), i32, P(qgpu_join_filter), P(vp)]
    lib.qgpu_plan_nested_loop_join.argtypes = [vp, vp, vp, i32, P(qgpu_join_filter), P(vp)]
    lib.qgpu_plan_cross_join.argtypes = [vp, vp, vp, P(vp)]
    lib.qgpu_plan_broadcast.argtypes = [vp, vp, i32, P(vp)]
    lib.qgpu_plan_broadcast_pruned.argtypes = [vp, vp, i32, vp, i32, P(vp)]
    lib.qgpu_plan_final_aggregate.argtypes = [vp, vp, P(i32), i32, P(i32), P(i32), i32, P(vp)]
    lib.qgpu_plan_schema.argtypes = [vp, vp]
    lib.qgpu_plan_execute.argtypes = [vp, vp]
    lib.qgpu_plan_execute_device.argtypes = [vp, P(vp), P(i64)]
    lib.qgpu_plan_last_stats.argtypes = [vp, P(ctypes.c_double), P(i64)]
    lib.qgpu_plan_strategy.argtypes = [vp]
    lib.qgpu_plan_strategy.restype = ctypes.c_char_p
    lib.qgpu_plan_free.argtypes = [vp]
    lib.qgpu_plan_free.restype = None
    lib.qgpu_table_hash_partition.argtypes = [vp, i32, i32, P(vp), P(i64)]
    lib.qgpu_table_column_device_buffer.argtypes = [vp, i32, P(vp), P(i64), P(i32)]
    lib.qgpu_release_cached_memory.argtypes = [vp]
    lib.qgpu_plan_sort.argtypes = [vp, vp, P(vp), P(i32), P(i32), i32, i64, P(vp)]
    lib.qgpu_plan_limit.argtypes = [vp, vp, i64, i64, P(vp)]
    lib.qgpu_plan_exchange_keystats.argtypes = [vp, P(i64), P(i32)]
    lib.qgpu_plan_exchange_sketch.argtypes = [vp, P(i64), P(vp), P(i64), P(i32)]
    lib.qgpu_plan_exchange_prepare.argtypes = [vp, vp, i32, i32, vp, P(i32), P(i32)]
    lib.qgpu_plan_exchange_scatter.argtypes = [vp, vp]
    lib.qgpu_plan_exchange_finish.argtypes = [vp, P(i32)]
    lib.qgpu_plan_state_bytes.argtypes = [vp, i32, P(i64)]
    lib.qgpu_plan_partial_state.argtypes = [vp, i64, i32, vp, i64]
    lib.qgpu_plan_execute_merged.argtypes = [vp, vp, i32, i32, vp]
    lib.qgpu_plan_execute_merged_device.argtypes = [vp, vp, i32, i32, P(vp)]
    lib.qgpu_plan_set_order_free.argtypes = [vp, i32]
    lib.qgpu_plan_execute_device_async.argtypes = [vp, P(vp)]
    lib.qgpu_table_wait.argtypes = [vp]
    lib.qgpu_comm_unique_id.argtypes = [vp, i64]
    lib.qgpu_comm_init.argtypes = [vp, vp, i32, i32]
    lib.qgpu_comm_init_local.argtypes = [P(vp), i32]
    lib.qgpu_comm_destroy.argtypes = [vp]
    lib.qgpu_comm_world.argtypes = [vp, P(i32), P(i32)]
    lib.qgpu_comm_all_gather.argtypes = [vp, vp, vp, i64]
    lib.qgpu_comm_all_to_all.argtypes = [vp, vp, P(i64), P(i64), vp, P(i64), P(i64)]
    lib.qgpu_comm_barrier.argtypes = [vp]
    lib.qgpu_plan_execute_sharded.argtypes = [vp, i64, i32, vp]
    lib.qgpu_plan_execute_sharded_device.argtypes = [vp, i64, i32, i32, P(vp)]
    _lib = lib
    return lib


def _addr(cdata) -> int:
    return int(_ffi.cast("uintptr_t", cdata))


class Context:
    """One qgpu_ctx == one GPU (one process per GPU)."""

    def __init__(self, device: Optional[int] = None):
        self.lib = load_library()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device
        h = ctypes.c_void_p()
        dev = (ctypes.c_int * 1)(device)
        rc = self.lib.qgpu_init(dev, 1, ctypes.byref(h))
        if rc != 0:
            raise QuriousError(rc, self.lib.qgpu_last_error(None).decode())
        self.handle = h

    def check(self, rc: int):
        if rc != 0:
            raise QuriousError(rc, self.lib.qgpu_last_error(self.handle).decode())

    def set_compat(self, name: str, value: bool):
        self.check(self.lib.qgpu_set_compat(self.handle, name.encode(), 1 if value else 0))

    def set_option(self, name: str, value: int):
        """qgpu_set_option: "ingest_threads", "ingest_host_narrow"."""
        self.check(self.lib.qgpu_set_option(self.handle, name.encode(), int(value)))

    def release_cached_memory(self):
        """Give the context's cached large device blocks back to the driver."""
        self.check(self.lib.qgpu_release_cached_memory(self.handle))

    def counter(self, name: str) -> int:
        """qgpu_counter: "alloc_bytes", "gather_bytes", "kernel_launches"."""
        return int(self.lib.qgpu_counter(self.handle, name.encode()))

    def kernel_launches(self) -> int:
        return int(self.lib.qgpu_kernel_launches(self.handle))

    def stream_handle(self) -> int:
        """cudaStream_t of the context's compute stream (for caller-side CUDA events)."""
        return int(self.lib.qgpu_ctx_stream(self.handle) or 0)

    def profile(self, on, min_blocks: int = 0):
        """on: bracket kernel launches with CUDA events; min_blocks >= 2: only launches of at least that many blocks."""
        self.check(self.lib.qgpu_profile_enable(self.handle, (max(int(min_blocks), 1) if on else 0)))

    def profile_report(self):
        """[(kernel, launches, total_ms, max_ms)] since the last report; device time from CUDA events."""
        cap = 1 << 16
        buf = ctypes.create_string_buffer(cap)
        need = int(self.lib.qgpu_profile_report(self.handle, buf, cap))
        if need < 0:
            raise QuriousError(1, self.lib.qgpu_last_error(self.handle).decode())
        out = []
        for line in buf.value.decode().splitlines():
            name, n, ms, mx = line.split("\t")
            out.append((name, int(n), float(ms), float(mx)))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.lib.qgpu_shutdown(self.handle)
            self.handle = None

    # ---- communicator (multi-GPU, include/qgpu.h "communicator") ------------------------------------
    def comm_unique_id(self) -> bytes:
        """128-byte NCCL unique id (rank 0 creates it, every rank passes it to comm_init)."""
        buf = ctypes.create_string_buffer(128)
        rc = self.lib.qgpu_comm_unique_id(buf, 128)
        if rc != 0:
            raise QuriousError(rc, self.lib.qgpu_last_error(None).decode())
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self.check(self.lib.qgpu_comm_init(self.handle, buf, rank, world))

    def comm_destroy(self):
        self.check(self.lib.qgpu_comm_destroy(self.handle))

    def comm_world(self):
        r, w = ctypes.c_int32(), ctypes.c_int32()
        self.lib.qgpu_comm_world(self.handle, ctypes.byref(r), ctypes.byref(w))
        return r.value, w.value

    def comm_barrier(self):
        self.check(self.lib.qgpu_comm_barrier(self.handle))

    def comm_all_gather(self, send_ptr: int, recv_ptr: int, bytes_per_rank: int):
        self.check(self.lib.qgpu_comm_all_gather(self.handle, send_ptr, recv_ptr, bytes_per_rank))

    def comm_all_to_all(self, send_ptr, send_off, send_bytes, recv_ptr, recv_off, recv_bytes):
        n = len(send_off)
        A = ctypes.c_int64 * n
        self.check(self.lib.qgpu_comm_all_to_all(self.handle, send_ptr, A(*send_off), A(*send_bytes), recv_ptr, A(*recv_off), A(*recv_bytes)))

    # ---- marshalling helpers ------------------------------------------------------------------
    def export_schema(self, schema: pa.Schema):
        c = _ffi.new("struct ArrowSchema*")
        schema._export_to_c(_addr(c))
        return c

    def parse_expr(self, expr) -> ctypes.c_void_p:
        global _parse_ctx
        prev, _parse_ctx = _parse_ctx, self           # SubQuery.to_ir builds its sub-plan in this context
        try:
            ir = expr.to_ir()
        finally:
            _parse_ctx = prev
        h = ctypes.c_void_p()
        self.check(self.lib.qgpu_expr_parse(self.handle, ir, len(ir), ctypes.byref(h)))
        return h


_default_ctx: Optional[Context] = None
_parse_ctx: Optional[Context] = None


def current_parse_context() -> Optional[Context]:
    return _parse_ctx


def comm_init_local(ctxs: Sequence["Context"]):
    """Wire several contexts of THIS process into one group without NCCL (ranks emulated on one GPU, tests)."""
    arr = (ctypes.c_void_p * len(ctxs))(*[c.handle for c in ctxs])
    rc = ctxs[0].lib.qgpu_comm_init_local(arr, len(ctxs))
    ctxs[0].check(rc)


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


class DeviceTable:
    """Owning wrapper of a qgpu_table handle (an HBM-resident MemoryTable)."""

    def __init__(self, ctx: Context, handle: ctypes.c_void_p, schema: pa.Schema):
        self.ctx = ctx
        self.handle = handle
        self.schema = schema

    @staticmethod
    def create(ctx: Context, schema: pa.Schema) -> "DeviceTable":
        h = ctypes.c_void_p()
        cs = ctx.export_schema(schema)
        try:
            ctx.check(ctx.lib.qgpu_table_create(ctx.handle, _addr(cs), ctypes.byref(h)))
        finally:
            if cs.release != _ffi.NULL:
                cs.release(cs)
        return DeviceTable(ctx, h, schema)

    def append(self, batch: pa.RecordBatch, upload_columns: Optional[Sequence[int]] = None):
        ca = _ffi.new("struct ArrowArray*")
        batch._export_to_c(_addr(ca))
        if upload_columns is None:
            rc = self.ctx.lib.qgpu_table_append(self.handle, _addr(ca), None, 0)
        else:
            arr = (ctypes.c_int32 * len(upload_columns))(*upload_columns)
            rc = self.ctx.lib.qgpu_table_append(self.handle, _addr(ca), arr, len(upload_columns))
        self.ctx.check(rc)

    def append_batches(self, batches: Sequence[pa.RecordBatch], upload_columns: Optional[Sequence[int]] = None) -> int:
        """Every batch through ONE FFI call (an Arrow C stream over the batches): the reference's tables arrive as tens of
        thousands of 1024-row batches (datasource/file/csv.rs:34-72).  -> batches appended"""
        reader = pa.RecordBatchReader.from_batches(self.schema, batches)
        cstream = _ffi.new("struct ArrowArrayStream*")
        reader._export_to_c(_addr(cstream))
        got = ctypes.c_int64()
        if upload_columns is None:
            rc = self.ctx.lib.qgpu_table_append_stream(self.handle, _addr(cstream), None, 0, ctypes.byref(got))
        else:
            arr = (ctypes.c_int32 * len(upload_columns))(*upload_columns)
            rc = self.ctx.lib.qgpu_table_append_stream(self.handle, _addr(cstream), arr, len(upload_columns), ctypes.byref(got))
        self.ctx.check(rc)
        return got.value

    def append_csv(self, source, has_header: bool = True, delimiter: str = ",", quote: Optional[str] = None,
                   escape: Optional[str] = None, upload_columns: Optional[Sequence[int]] = None) -> int:
        """`read_csv(path, CsvReadOptions{has_header, delimiter, quote, escape})` (datasource/file/csv.rs:16-72) into this
        table's declared schema, parsed on the GPU.  source: a path (str) or the file's bytes.  -> rows appended"""
        opt = qgpu_csv_options(1 if has_header else 0, ord(delimiter), ord(quote) if quote else 0, ord(escape) if escape else 0)
        got = ctypes.c_int64()
        arr, n = (None, 0) if upload_columns is None else ((ctypes.c_int32 * len(upload_columns))(*upload_columns), len(upload_columns))
        if isinstance(source, (bytes, bytearray, memoryview)):
            b = bytes(source)
            rc = self.ctx.lib.qgpu_table_append_csv(self.handle, b, len(b), ctypes.byref(opt), arr, n, ctypes.byref(got))
        else:
            rc = self.ctx.lib.qgpu_table_append_csv_file(self.handle, str(source).encode(), ctypes.byref(opt), arr, n, ctypes.byref(got))
        self.ctx.check(rc)
        return got.value

    def flush(self):
        """Upload the retained host batches now (otherwise they are uploaded when the table is first used)."""
        self.ctx.check(self.ctx.lib.qgpu_table_flush(self.handle))

    def append_device_struct(self, array_addr: int):
        """`array_addr`: address of a struct ArrowArray whose buffers are device pointers."""
        self.ctx.check(self.ctx.lib.qgpu_table_append_device(self.handle, array_addr))

    @property
    def num_rows(self) -> int:
        n = int(self.ctx.lib.qgpu_table_num_rows(self.handle))
        if n < 0:
            raise QuriousError(1, self.ctx.lib.qgpu_last_error(self.ctx.handle).decode())
        return n

    def wait(self) -> "DeviceTable":
        """Result of an asynchronous execute: wait for its metadata; errors of the producing kernels are raised here."""
        self.ctx.check(self.ctx.lib.qgpu_table_wait(self.handle))
        return self

    @property
    def num_batches(self) -> int:
        return int(self.ctx.lib.qgpu_table_num_batches(self.handle))

    def column_bytes(self, col: int) -> int:
        r = int(self.ctx.lib.qgpu_table_column_bytes(self.handle, col))
        if r < 0:
            raise QuriousError(1, self.ctx.lib.qgpu_last_error(self.ctx.handle).decode())
        return r

    def hash_partition(self, key_col: int, n_parts: int):
        """-> (DeviceTable with rows grouped by partition, [n_parts + 1] row offsets)."""
        out = ctypes.c_void_p()
        offs = (ctypes.c_int64 * (n_parts + 1))()
        self.ctx.check(self.ctx.lib.qgpu_table_hash_partition(self.handle, key_col, n_parts, ctypes.byref(out), offs))
        return DeviceTable(self.ctx, out, self.schema), list(offs)

    def column_device_buffer(self, col: int):
        """-> (device pointer, bytes, value width) of a fixed-width column's value buffer."""
        p, n, w = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int32()
        self.ctx.check(self.ctx.lib.qgpu_table_column_device_buffer(self.handle, col, ctypes.byref(p), ctypes.byref(n),
                                                                   ctypes.byref(w)))
        return int(p.value or 0), n.value, w.value

    def to_batch(self) -> pa.RecordBatch:
        ca = _ffi.new("struct ArrowArray*")
        cs = _ffi.new("struct ArrowSchema*")
        self.ctx.check(self.ctx.lib.qgpu_table_export(self.handle, _addr(ca), _addr(cs)))
        return pa.RecordBatch._import_from_c(_addr(ca), _addr(cs))

    def free(self):
        if self.handle:
            self.ctx.lib.qgpu_table_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            if self.ctx.handle:
                self.free()
        except Exception:
            pass


def read_stream(ctx: Context, stream_cdata) -> List[pa.RecordBatch]:
    reader = pa.RecordBatchReader._import_from_c(_addr(stream_cdata))
    return [b for b in reader]


def new_stream():
    return _ffi.new("struct ArrowArrayStream*")


addr = _addr
