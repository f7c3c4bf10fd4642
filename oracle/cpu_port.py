"""ctypes wrapper of oracle/qref_cpu.cpp (the single-threaded C++ "port" of qurious's CPU steps).

TEST/BENCH INFRASTRUCTURE ONLY (see the header of qref_cpu.cpp).  Used by bench.py's `cpu_baseline`
and `--impl reference` arms and, in tests/, as a second checker next to oracle/qref.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Tuple

import numpy as np
import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libqref_cpu.so")
_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-C", _HERE])
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.qcpu_q1.restype = ctypes.c_int64
        _lib.qcpu_groupby.restype = ctypes.c_int64
        _lib.qcpu_q3.restype = ctypes.c_int64
    return _lib


def _buf(arr: pa.Array, i: int) -> Tuple[int, object]:
    b = arr.buffers()[i]
    return b.address, b


def _fixed(arr: pa.Array, width: int) -> int:
    assert arr.null_count == 0
    return arr.buffers()[1].address + arr.offset * width


def _one_chunk(table_batches: List[pa.RecordBatch], name: str) -> pa.Array:
    t = pa.Table.from_batches(table_batches).column(name)
    return t.chunk(0) if t.num_chunks == 1 else pa.concat_arrays(t.chunks)


def _i128(lo_hi: np.ndarray) -> int:
    lo, hi = int(lo_hi[0]), int(lo_hi[1])
    v = (hi << 64) | lo
    return v - (1 << 128) if v >= (1 << 127) else v


def q6(batches: List[pa.RecordBatch], batch_rows: int = 1024) -> Dict[str, int]:
    """-> {"revenue_raw": unscaled Decimal128(31,4) SUM (or None), "rows": rows passing the filter}."""
    lib = load()
    cols = {n: _one_chunk(batches, n) for n in ("l_shipdate", "l_discount", "l_quantity", "l_extendedprice")}
    n = len(cols["l_shipdate"])
    out = np.zeros(2, dtype=np.uint64)
    rows = ctypes.c_int64()
    lib.qcpu_q6(ctypes.c_int64(n), ctypes.c_int64(batch_rows), ctypes.c_void_p(_fixed(cols["l_shipdate"], 4)),
                ctypes.c_void_p(_fixed(cols["l_discount"], 16)), ctypes.c_void_p(_fixed(cols["l_quantity"], 16)),
                ctypes.c_void_p(_fixed(cols["l_extendedprice"], 16)), b"1994-01-01", b"1995-01-01",
                ctypes.c_double(0.049999999999999996), ctypes.c_double(0.06999999999999999), ctypes.c_int64(24),
                out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(rows))
    return {"revenue_raw": _i128(out) if rows.value > 0 else None, "rows": rows.value}


def q1(batches: List[pa.RecordBatch], batch_rows: int = 1024, max_groups: int = 64) -> List[tuple]:
    """-> rows (rf, ls, sum_qty, sum_base_price, sum_disc_price, sum_charge, avg_qty, avg_price, avg_disc, count)
    with decimals as unscaled ints, in first-occurrence order."""
    lib = load()
    names = ("l_shipdate", "l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax")
    c = {n: _one_chunk(batches, n) for n in names}
    n = len(c["l_shipdate"])
    for s in ("l_returnflag", "l_linestatus"):
        assert c[s].offset == 0 and c[s].null_count == 0
    rf, ls = c["l_returnflag"].buffers(), c["l_linestatus"].buffers()
    out_rf = ctypes.create_string_buffer(max_groups)
    out_ls = ctypes.create_string_buffer(max_groups)
    sums = np.zeros((max_groups, 4, 2), dtype=np.uint64)
    avgs = np.zeros((max_groups, 3, 2), dtype=np.uint64)
    cnt = np.zeros(max_groups, dtype=np.int64)
    vp = ctypes.c_void_p
    g = lib.qcpu_q1(ctypes.c_int64(n), ctypes.c_int64(batch_rows), vp(_fixed(c["l_shipdate"], 4)), vp(rf[1].address),
                    vp(rf[2].address), vp(ls[1].address), vp(ls[2].address), vp(_fixed(c["l_quantity"], 16)),
                    vp(_fixed(c["l_extendedprice"], 16)), vp(_fixed(c["l_discount"], 16)), vp(_fixed(c["l_tax"], 16)),
                    b"1998-09-02", ctypes.c_int64(max_groups), out_rf, out_ls, sums.ctypes.data_as(vp),
                    avgs.ctypes.data_as(vp), cnt.ctypes.data_as(vp))
    rows = []
    for i in range(g):
        rows.append((out_rf.raw[i:i + 1].decode(), out_ls.raw[i:i + 1].decode(),
                     *[_i128(sums[i, a]) for a in range(4)], *[_i128(avgs[i, a]) for a in range(3)], int(cnt[i])))
    return rows


def groupby(batches: List[pa.RecordBatch], batch_rows: int = 1024, max_groups: int = None) -> List[tuple]:
    """config 4: -> rows (k, sum_v, count_v, min_v, max_v, avg_f) in first-occurrence order."""
    lib = load()
    c = {n: _one_chunk(batches, n) for n in ("k", "v", "f")}
    n = len(c["k"])
    mg = n if max_groups is None else max_groups
    ok, os_, oc, omn, omx = (np.zeros(max(mg, 1), dtype=np.int64) for _ in range(5))
    oa = np.zeros(max(mg, 1), dtype=np.float64)
    vp = ctypes.c_void_p
    g = lib.qcpu_groupby(ctypes.c_int64(n), ctypes.c_int64(batch_rows), vp(_fixed(c["k"], 8)), vp(_fixed(c["v"], 8)),
                         vp(_fixed(c["f"], 8)), ctypes.c_int64(mg), *[a.ctypes.data_as(vp) for a in (ok, os_, oc, omn, omx, oa)])
    return [(int(ok[i]), int(os_[i]), int(oc[i]), int(omn[i]), int(omx[i]), float(oa[i])) for i in range(g)]


def q3(customer: List[pa.RecordBatch], orders: List[pa.RecordBatch], lineitem: List[pa.RecordBatch], batch_rows: int = 1024,
       max_groups: int = None) -> List[tuple]:
    """-> rows (l_orderkey, revenue_raw [unscaled Decimal128(38,4)], o_orderdate [days], o_shippriority), first-occurrence order."""
    lib = load()
    cc = {n: _one_chunk(customer, n) for n in ("c_custkey", "c_mktsegment")}
    oc = {n: _one_chunk(orders, n) for n in ("o_orderkey", "o_custkey", "o_orderdate", "o_shippriority")}
    lc = {n: _one_chunk(lineitem, n) for n in ("l_orderkey", "l_shipdate", "l_extendedprice", "l_discount")}
    assert cc["c_mktsegment"].offset == 0 and cc["c_mktsegment"].null_count == 0
    seg = cc["c_mktsegment"].buffers()
    n_l = len(lc["l_orderkey"])
    mg = n_l if max_groups is None else max_groups
    o_key, o_prio = np.zeros(max(mg, 1), dtype=np.int64), np.zeros(max(mg, 1), dtype=np.int64)
    o_date = np.zeros(max(mg, 1), dtype=np.int32)
    o_rev = np.zeros((max(mg, 1), 2), dtype=np.uint64)
    vp, i64 = ctypes.c_void_p, ctypes.c_int64
    g = lib.qcpu_q3(i64(len(cc["c_custkey"])), vp(_fixed(cc["c_custkey"], 8)), vp(seg[1].address), vp(seg[2].address),
                    i64(len(oc["o_orderkey"])), vp(_fixed(oc["o_orderkey"], 8)), vp(_fixed(oc["o_custkey"], 8)),
                    vp(_fixed(oc["o_orderdate"], 4)), vp(_fixed(oc["o_shippriority"], 8)),
                    i64(n_l), vp(_fixed(lc["l_orderkey"], 8)), vp(_fixed(lc["l_shipdate"], 4)), vp(_fixed(lc["l_extendedprice"], 16)),
                    vp(_fixed(lc["l_discount"], 16)), i64(batch_rows), b"BUILDING", b"1995-03-15", i64(mg),
                    o_key.ctypes.data_as(vp), o_rev.ctypes.data_as(vp), o_date.ctypes.data_as(vp), o_prio.ctypes.data_as(vp))
    return [(int(o_key[i]), _i128(o_rev[i]), int(o_date[i]), int(o_prio[i])) for i in range(g)]
