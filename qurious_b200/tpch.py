"""Synthetic TPC-H-shaped tables and the physical plans qurious builds for Q1/Q6/Q3.

Data: deterministic, counter-based (value = hash(seed, column, row)), integer-only arithmetic on
torch int64 tensors so the very same code generates identical tables on the host (tests, oracle,
CPU baseline) and directly in HBM (bench.py's device-resident leg).  No network, no dbgen: the
VALUES therefore differ from the reference's SF0.01 goldens (tests/tpch/q*.slt); types, schemas
and plan shapes are the reference's (tests/tpch/create_tables.slt:41-84, planner/sql.rs:1439-1476:
BIGINT/INTEGER -> Int64, DECIMAL(15,2) -> Decimal128(15,2), DATE -> Date32, VARCHAR -> Utf8).

Plans: exactly the operator trees of SURVEY.md 3.2-3.4 (after the reference's 8 optimizer rules):
pushed-down WHERE inside Scan, literal casts kept as CastExpr(Literal) (folded once by the library),
build side = left child.  Sort/Limit are outside the hot path and are not part of
these plans; callers order the (small) result themselves.
"""
from __future__ import annotations

import datetime
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import pyarrow as pa
import torch

from .datatypes import JoinType, Operator, ScalarValue
from .physical.expr import (AvgAggregateExpr, BinaryExpr, CastExpr, Column, CountAggregateExpr, Literal,
                            SumAggregateExpr, avg_return_type)
from .physical.plan import HashAggregate, HashJoinExec, MemoryTable, NoGroupingAggregate, Projection, Scan

SEED = 20240601
DEC = pa.decimal128(15, 2)
EPOCH = datetime.date(1970, 1, 1)


def days(s: str) -> int:
    return (datetime.date.fromisoformat(s) - EPOCH).days


LINEITEM_ROWS = {1.0: 6001215, 10.0: 59986052, 100.0: 600037902}
SEGMENTS = ["AUTOMOBILE", "BUILDING", "FURNITURE", "HOUSEHOLD", "MACHINERY"]
SHIPINSTRUCT = ["DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN"]
SHIPMODE = ["REG AIR", "AIR", "RAIL", "SHIP", "TRUCK", "MAIL", "FOB"]
PRIORITY = ["1-URGENT", "2-HIGH", "3-MEDIUM", "4-NOT SPECIFIED", "5-LOW"]
_WORDS = ["furiously", "sly", "careful", "blithe", "quick", "fluffy", "slow", "quiet", "ruthless", "thin", "close",
          "dogged", "daring", "bold", "ironic", "final", "pending", "regular", "express", "special", "deposits",
          "requests", "accounts", "packages", "foxes", "ideas", "theodolites", "pinto beans", "instructions"]
COMMENTS = [" ".join(_WORDS[(i * 7 + j * 3) % len(_WORDS)] for j in range(2 + i % 4)) for i in range(64)]

LINEITEM_SCHEMA = pa.schema([
    ("l_orderkey", pa.int64()), ("l_partkey", pa.int64()), ("l_suppkey", pa.int64()), ("l_linenumber", pa.int64()),
    ("l_quantity", DEC), ("l_extendedprice", DEC), ("l_discount", DEC), ("l_tax", DEC),
    ("l_returnflag", pa.string()), ("l_linestatus", pa.string()), ("l_shipdate", pa.date32()),
    ("l_commitdate", pa.date32()), ("l_receiptdate", pa.date32()), ("l_shipinstruct", pa.string()),
    ("l_shipmode", pa.string()), ("l_comment", pa.string()), ("l_rev", pa.string())])
ORDERS_SCHEMA = pa.schema([
    ("o_orderkey", pa.int64()), ("o_custkey", pa.int64()), ("o_orderstatus", pa.string()), ("o_totalprice", DEC),
    ("o_orderdate", pa.date32()), ("o_orderpriority", pa.string()), ("o_clerk", pa.string()),
    ("o_shippriority", pa.int64()), ("o_comment", pa.string()), ("o_rev", pa.string())])
CUSTOMER_SCHEMA = pa.schema([
    ("c_custkey", pa.int64()), ("c_name", pa.string()), ("c_address", pa.string()), ("c_nationkey", pa.int64()),
    ("c_phone", pa.string()), ("c_acctbal", DEC), ("c_mktsegment", pa.string()), ("c_comment", pa.string()),
    ("c_rev", pa.string())])


# ------------------------------------------------------------------------------------------------
# counter-based integer RNG on int64 tensors (wrapping multiply, logical shifts emulated)
# ------------------------------------------------------------------------------------------------
def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    return (x >> k) & ((1 << (64 - k)) - 1)


def _mix(x: torch.Tensor) -> torch.Tensor:
    x = x ^ _lsr(x, 30)
    x = x * (-4658895280553007687)       # 0xBF58476D1CE4E5B9 as int64
    x = x ^ _lsr(x, 27)
    x = x * (-7723592293110705685)       # 0x94D049BB133111EB as int64
    return x ^ _lsr(x, 31)


def rnd(col_id: int, idx: torch.Tensor) -> torch.Tensor:
    """Non-negative pseudo-random int64 (62 bits) for (SEED, column id, row index)."""
    salt = (SEED * 1000003 + col_id * 7919) * 2654435761 % (1 << 62)
    return _lsr(_mix(idx * (-7046029254386353131) + salt), 2)


def uniform(col_id: int, idx: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
    return lo + rnd(col_id, idx) % (hi - lo + 1)


# ------------------------------------------------------------------------------------------------
# column generation (torch tensors; device = "cpu" or "cuda")
# ------------------------------------------------------------------------------------------------
@dataclass
class RawTable:
    name: str
    schema: pa.Schema
    rows: int
    cols: Dict[str, torch.Tensor]          # numeric columns: int64 (decimals = raw unscaled value), int32 dates
    codes: Dict[str, torch.Tensor]         # string columns as vocabulary codes
    vocab: Dict[str, List[str]]


def n_orders(sf: float) -> int:
    return max(int(round(1_500_000 * sf)), 8)


def n_customers(sf: float) -> int:
    return max(int(round(150_000 * sf)), 8)


def n_suppliers(sf: float) -> int:
    return max(int(10_000 * sf), 100)


def n_parts(sf: float) -> int:
    return max(int(200_000 * sf), 1000)


def supplier_of(partkey: torch.Tensor, j: torch.Tensor, sf: float) -> torch.Tensor:
    """The j-th (0..3) supplier of a part: partsupp holds exactly these four (partkey, suppkey) pairs per part."""
    s = n_suppliers(sf)
    return ((partkey - 1) + j * (s // 4)) % s + 1


def n_lineitems(sf: float) -> int:
    return LINEITEM_ROWS.get(float(sf), max(int(round(6_001_215 * sf)), n_orders(sf)))


def _orderkey(i: torch.Tensor) -> torch.Tensor:
    return (i // 8) * 32 + (i % 8) + 1   # dbgen's sparse keys: first 8 of every 32


def gen_customer(sf: float, device="cpu", columns: Optional[Sequence[str]] = None) -> RawTable:
    n = n_customers(sf)
    i = torch.arange(n, dtype=torch.int64, device=device)
    want = set(columns) if columns is not None else set(CUSTOMER_SCHEMA.names)
    cols, codes, vocab = {}, {}, {}
    if "c_custkey" in want:
        cols["c_custkey"] = i + 1
    if "c_nationkey" in want:
        cols["c_nationkey"] = uniform(32, i, 0, 24)
    if "c_acctbal" in want:
        cols["c_acctbal"] = uniform(33, i, -99999, 999999)
    for k, (cid, voc) in {"c_mktsegment": (34, SEGMENTS), "c_comment": (35, COMMENTS), "c_address": (36, COMMENTS),
                          "c_phone": (37, [f"{10 + j}-{100 + j}-{200 + j}-{1000 + j}" for j in range(25)]),
                          "c_name": (38, [f"Customer#{j:09d}" for j in range(1000)]), "c_rev": (39, [""])}.items():
        if k in want:
            codes[k] = uniform(cid, i, 0, len(voc) - 1)
            vocab[k] = voc
    schema = pa.schema([f for f in CUSTOMER_SCHEMA if f.name in want])
    return RawTable("customer", schema, n, cols, codes, vocab)


def _order_dates(i: torch.Tensor) -> torch.Tensor:
    return uniform(20, i, days("1992-01-01"), days("1998-08-02"))


def gen_orders(sf: float, device="cpu", columns: Optional[Sequence[str]] = None, row_range=None) -> RawTable:
    """row_range=(lo, hi): only that contiguous slice of orders (multi-GPU: the build side of Q3 is sharded too)."""
    n = n_orders(sf)
    nc = n_customers(sf)
    lo, hi = row_range if row_range is not None else (0, n)
    i = torch.arange(lo, hi, dtype=torch.int64, device=device)
    n = hi - lo
    want = set(columns) if columns is not None else set(ORDERS_SCHEMA.names)
    cols, codes, vocab = {}, {}, {}
    if "o_orderkey" in want:
        cols["o_orderkey"] = _orderkey(i)
    if "o_custkey" in want:
        k = rnd(21, i) % (nc - nc // 3)
        cols["o_custkey"] = k + k // 2 + 1      # custkeys not divisible by 3
    if "o_totalprice" in want:
        cols["o_totalprice"] = uniform(22, i, 90000, 50000000)
    if "o_orderdate" in want:
        cols["o_orderdate"] = _order_dates(i).to(torch.int32)
    if "o_shippriority" in want:
        cols["o_shippriority"] = torch.zeros(n, dtype=torch.int64, device=device)
    for k, (cid, voc) in {"o_orderstatus": (23, ["O", "F", "P"]), "o_orderpriority": (24, PRIORITY),
                          "o_clerk": (25, [f"Clerk#{j:09d}" for j in range(1000)]), "o_comment": (26, COMMENTS),
                          "o_rev": (27, [""])}.items():
        if k in want:
            codes[k] = uniform(cid, i, 0, len(voc) - 1)
            vocab[k] = voc
    schema = pa.schema([f for f in ORDERS_SCHEMA if f.name in want])
    return RawTable("orders", schema, n, cols, codes, vocab)


def _lines_per_order(sf: float, device) -> torch.Tensor:
    """1..7 lines per order, adjusted deterministically so that the total is exactly n_lineitems(sf)."""
    no, target = n_orders(sf), n_lineitems(sf)
    o = torch.arange(no, dtype=torch.int64, device=device)
    L = 1 + rnd(1, o) % 7
    diff = target - int(L.sum().item())
    if diff > 0:
        can = (L < 7).to(torch.int64)
        for _ in range(8):  # a few rounds are enough: each round can add up to #orders lines
            rank = torch.cumsum(can, 0) - can
            add = can * (rank < diff).to(torch.int64)
            L = L + add
            diff -= int(add.sum().item())
            if diff <= 0:
                break
            can = torch.ones_like(L)  # beyond 7 lines only when the target cannot be met otherwise
    elif diff < 0:
        need = -diff
        for _ in range(8):
            can = (L > 1).to(torch.int64)
            rank = torch.cumsum(can, 0) - can
            sub = can * (rank < need).to(torch.int64)
            L = L - sub
            need -= int(sub.sum().item())
            if need <= 0:
                break
    return L


def gen_lineitem(sf: float, device="cpu", columns: Optional[Sequence[str]] = None, row_range=None) -> RawTable:
    """row_range=(lo, hi): generate only that contiguous slice of lineitem (multi-GPU row-range sharding);
    every value depends only on (order index, line number) so shards are consistent."""
    no = n_orders(sf)
    L = _lines_per_order(sf, device)
    n_total = int(L.sum().item())
    o_idx = torch.repeat_interleave(torch.arange(no, dtype=torch.int64, device=device), L)
    starts = torch.cumsum(L, 0) - L
    r = torch.arange(n_total, dtype=torch.int64, device=device)
    line = r - starts[o_idx]
    if row_range is not None:
        lo, hi = row_range
        o_idx, line, r = o_idx[lo:hi], line[lo:hi], r[lo:hi]
    n = int(r.numel())
    key = o_idx * 8 + line                      # unique per (order, line): the RNG counter
    want = set(columns) if columns is not None else set(LINEITEM_SCHEMA.names)
    cols, codes, vocab = {}, {}, {}
    odate = _order_dates(o_idx)
    ship = odate + uniform(2, key, 1, 121)
    receipt = ship + uniform(4, key, 1, 30)
    qty = uniform(5, key, 1, 50)
    pk = uniform(6, key, 1, n_parts(sf))
    retail = 90000 + (pk // 10) % 20001 + 100 * (pk % 1000)
    cutoff = days("1995-06-17")
    if "l_orderkey" in want:
        cols["l_orderkey"] = _orderkey(o_idx)
    if "l_partkey" in want:
        cols["l_partkey"] = pk
    if "l_suppkey" in want:
        cols["l_suppkey"] = supplier_of(pk, rnd(7, key) % 4, sf)       # one of the part's four partsupp rows (Q9 joins them)
    if "l_linenumber" in want:
        cols["l_linenumber"] = line + 1
    if "l_quantity" in want:
        cols["l_quantity"] = qty * 100
    if "l_extendedprice" in want:
        cols["l_extendedprice"] = qty * retail
    if "l_discount" in want:
        cols["l_discount"] = uniform(8, key, 0, 10)
    if "l_tax" in want:
        cols["l_tax"] = uniform(9, key, 0, 8)
    if "l_shipdate" in want:
        cols["l_shipdate"] = ship.to(torch.int32)
    if "l_commitdate" in want:
        cols["l_commitdate"] = (odate + uniform(3, key, 30, 90)).to(torch.int32)
    if "l_receiptdate" in want:
        cols["l_receiptdate"] = receipt.to(torch.int32)
    if "l_returnflag" in want:
        ra = rnd(10, key) % 2                    # 0 -> 'R', 1 -> 'A'
        codes["l_returnflag"] = torch.where(receipt <= cutoff, ra, torch.full_like(ra, 2))
        vocab["l_returnflag"] = ["R", "A", "N"]
    if "l_linestatus" in want:
        codes["l_linestatus"] = (ship > cutoff).to(torch.int64)
        vocab["l_linestatus"] = ["F", "O"]
    for k, (cid, voc) in {"l_shipinstruct": (11, SHIPINSTRUCT), "l_shipmode": (12, SHIPMODE), "l_comment": (13, COMMENTS),
                          "l_rev": (14, [""])}.items():
        if k in want:
            codes[k] = uniform(cid, key, 0, len(voc) - 1)
            vocab[k] = voc
    schema = pa.schema([f for f in LINEITEM_SCHEMA if f.name in want])
    return RawTable("lineitem", schema, n, cols, codes, vocab)


# ------------------------------------------------------------------------------------------------
# RawTable -> Arrow (host)
# ------------------------------------------------------------------------------------------------
def _decimal_array(raw: np.ndarray, dt: pa.DataType) -> pa.Array:
    n = len(raw)
    buf = np.empty((n, 2), dtype=np.int64)
    buf[:, 0] = raw
    buf[:, 1] = raw >> 63
    return pa.Array.from_buffers(dt, n, [None, pa.py_buffer(buf.tobytes())], null_count=0)


def to_arrow(t: RawTable, batch_rows: Optional[int] = None) -> List[pa.RecordBatch]:
    arrays = []
    for f in t.schema:
        if f.name in t.cols:
            a = t.cols[f.name].cpu().numpy()
            if pa.types.is_decimal(f.type):
                arrays.append(_decimal_array(a.astype(np.int64), f.type))
            elif f.type == pa.date32():
                arrays.append(pa.array(a.astype(np.int32), type=pa.int32()).cast(pa.date32()))
            else:
                arrays.append(pa.array(a, type=f.type))
        else:
            code = t.codes[f.name].cpu().numpy().astype(np.int32)
            d = pa.DictionaryArray.from_arrays(pa.array(code), pa.array(t.vocab[f.name], type=pa.string()))
            arrays.append(d.cast(pa.string()))
    full = pa.record_batch(arrays, schema=t.schema)
    if batch_rows is None or batch_rows >= t.rows:
        return [full]
    return [full.slice(o, min(batch_rows, t.rows - o)) for o in range(0, t.rows, batch_rows)]


# ------------------------------------------------------------------------------------------------
# RawTable -> HBM-resident table (no host round trip): Arrow-layout device buffers handed to the
# library through qgpu_table_append_device (Arrow C Device Data Interface convention).
# ------------------------------------------------------------------------------------------------
def _device_string_buffers(codes: torch.Tensor, vocab: List[str]):
    dev = codes.device
    enc = [v.encode() for v in vocab]
    maxlen = max(max(len(e) for e in enc), 1)
    lens = torch.tensor([len(e) for e in enc], dtype=torch.int64, device=dev)
    table = torch.zeros((len(enc), maxlen), dtype=torch.uint8, device=dev)
    for i, e in enumerate(enc):
        if e:
            table[i, :len(e)] = torch.tensor(list(e), dtype=torch.uint8, device=dev)
    row_len = lens[codes]
    offsets = torch.zeros(codes.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(row_len, 0, out=offsets[1:])
    total = int(offsets[-1].item())
    if total >= (1 << 31):
        raise ValueError("Utf8 column exceeds 2 GiB")
    if bool((row_len == 1).all().item()):
        data = table[codes, 0].contiguous()
    else:
        row_of = torch.repeat_interleave(torch.arange(codes.numel(), device=dev), row_len)
        pos = torch.arange(total, device=dev) - offsets[row_of]
        data = table[codes[row_of], pos].contiguous()
    return offsets.to(torch.int32), data


def to_device_table(ctx, t: RawTable, keep_arrow_layout: bool = False):
    """Build an HBM-resident `MemoryTable` from a RawTable generated on `cuda` (bench.py device leg).
    Decimals are handed over in Arrow layout (16 B little-endian two's complement); the library narrows
    them itself exactly as it does for host batches."""
    from pyarrow.cffi import ffi
    from . import _lib
    n = t.rows
    keep = []
    children = []
    for f in t.schema:
        bufs: List[int] = [0]
        if f.name in t.cols:
            x = t.cols[f.name]
            if pa.types.is_decimal(f.type):
                x = torch.stack([x, x >> 63], dim=1).contiguous()
            elif f.type == pa.date32():
                x = x.to(torch.int32).contiguous()
            else:
                x = x.contiguous()
            keep.append(x)
            bufs.append(x.data_ptr())
        else:
            off, data = _device_string_buffers(t.codes[f.name], t.vocab[f.name])
            keep += [off, data]
            bufs += [off.data_ptr(), data.data_ptr() if data.numel() else 0]
        cb = ffi.new("const void*[]", [ffi.cast("void*", b) for b in bufs])
        ca = ffi.new("struct ArrowArray*")
        ca.length, ca.null_count, ca.offset = n, 0, 0
        ca.n_buffers, ca.n_children = len(bufs), 0
        ca.buffers = cb
        ca.release = ffi.NULL
        keep += [cb, ca]
        children.append(ca)
    top = ffi.new("struct ArrowArray*")
    kids = ffi.new("struct ArrowArray*[]", children)
    topbufs = ffi.new("const void*[]", [ffi.NULL])
    top.length, top.null_count, top.offset = n, 0, 0
    top.n_buffers, top.n_children = 1, len(children)
    top.buffers, top.children, top.release = topbufs, kids, ffi.NULL
    torch.cuda.synchronize()                     # generator kernels ran on torch's stream
    dev = _lib.DeviceTable.create(ctx, t.schema)
    dev.append_device_struct(_lib.addr(top))
    dev.column_bytes(0)                          # consolidates + synchronises the library's D2D copies
    del keep, kids, topbufs
    return MemoryTable.from_device_table(dev)


@dataclass
class Database:
    sf: float
    customer: MemoryTable
    orders: MemoryTable
    lineitem: MemoryTable


Q1_COLUMNS = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]
Q6_COLUMNS = ["l_quantity", "l_extendedprice", "l_discount", "l_shipdate"]
Q3_COLUMNS = {"customer": ["c_custkey", "c_mktsegment"],
              "orders": ["o_orderkey", "o_custkey", "o_orderdate", "o_shippriority"],
              "lineitem": ["l_orderkey", "l_extendedprice", "l_discount", "l_shipdate"]}


def generate(sf: float, batch_rows: Optional[int] = 1024, columns: Optional[Dict[str, Sequence[str]]] = None) -> Database:
    """Host tables (pyarrow RecordBatches of `batch_rows` rows -- the reference's CSV reader yields
    1024-row batches, datasource/file/csv.rs:63-66)."""
    columns = columns or {}
    c = gen_customer(sf, columns=columns.get("customer"))
    o = gen_orders(sf, columns=columns.get("orders"))
    l = gen_lineitem(sf, columns=columns.get("lineitem"))
    mk = lambda t: MemoryTable.try_new(t.schema, to_arrow(t, batch_rows))  # noqa: E731
    return Database(sf, mk(c), mk(o), mk(l))


# ------------------------------------------------------------------------------------------------
# plans
# ------------------------------------------------------------------------------------------------
def _col(schema: pa.Schema, name: str) -> Column:
    return Column(name, schema.get_field_index(name))


def _date(s: str) -> CastExpr:
    return CastExpr(Literal(ScalarValue.Utf8(s)), pa.date32())  # planner/sql.rs:1009-1012


def _one20() -> CastExpr:
    return CastExpr(Literal(ScalarValue.Int64(1)), pa.decimal128(20, 0))  # utils/type_coercion.rs:145-164


def _b(l, op: Operator, r) -> BinaryExpr:
    return BinaryExpr(l, op, r)


def q6_plan(db: Database) -> Projection:
    """SURVEY 3.2 (qurious/tests/tpch/q6.slt:1-12)."""
    s = db.lineitem.schema
    O = Operator
    ship, disc, qty, price = (_col(s, n) for n in ("l_shipdate", "l_discount", "l_quantity", "l_extendedprice"))
    pred = _b(_b(_b(_b(_b(ship, O.GtEq, _date("1994-01-01")), O.And, _b(ship, O.Lt, _date("1995-01-01"))), O.And,
                    _b(disc, O.GtEq, CastExpr(Literal(ScalarValue.Float64(0.049999999999999996)), DEC))), O.And,
                 _b(disc, O.LtEq, CastExpr(Literal(ScalarValue.Float64(0.06999999999999999)), DEC))), O.And,
              _b(qty, O.Lt, CastExpr(Literal(ScalarValue.Int64(24)), DEC)))
    rt = pa.decimal128(31, 4)
    agg = NoGroupingAggregate(pa.schema([("SUM(l_extendedprice * l_discount)", rt)]),
                              Scan(s, db.lineitem, None, pred), [SumAggregateExpr(_b(price, O.Mul, disc), rt)])
    return Projection(pa.schema([("revenue", rt)]), agg, [Column("revenue", 0)])


def q1_plan(db: Database) -> Projection:
    """SURVEY 3.3 (qurious/tests/tpch/q1.slt:1-22), without the final Sort."""
    s = db.lineitem.schema
    O = Operator
    qty, price, disc, tax, rf, ls, ship = (_col(s, n) for n in ("l_quantity", "l_extendedprice", "l_discount", "l_tax",
                                                                 "l_returnflag", "l_linestatus", "l_shipdate"))
    pred = _b(ship, O.LtEq, _date("1998-09-02"))
    disc_price = _b(price, O.Mul, _b(_one20(), O.Sub, disc))                     # Decimal128(38,4)
    charge = _b(disc_price, O.Mul, _b(_one20(), O.Add, tax))                     # Decimal128(38,6)
    t384, t386, avg_t = pa.decimal128(38, 4), pa.decimal128(38, 6), avg_return_type(DEC)
    aggs = [SumAggregateExpr(qty, DEC), SumAggregateExpr(price, DEC), SumAggregateExpr(disc_price, t384),
            SumAggregateExpr(charge, t386), AvgAggregateExpr(qty, DEC, avg_t), AvgAggregateExpr(price, DEC, avg_t),
            AvgAggregateExpr(disc, DEC, avg_t), CountAggregateExpr(Literal(ScalarValue.Int64(1)))]
    names = ["sum_qty", "sum_base_price", "sum_disc_price", "sum_charge", "avg_qty", "avg_price", "avg_disc", "count_order"]
    types = [DEC, DEC, t384, t386, avg_t, avg_t, avg_t, pa.int64()]
    schema = pa.schema([("l_returnflag", pa.string()), ("l_linestatus", pa.string())] + list(zip(names, types)))
    agg = HashAggregate(schema, Scan(s, db.lineitem, None, pred), [rf, ls], aggs)
    return Projection(schema, agg, [Column(f.name, i) for i, f in enumerate(schema)])


def q3_plan(db: Database) -> Projection:
    """SURVEY 3.4 (qurious/tests/tpch/q3.slt:1-24), without Sort/Limit: all groups are returned."""
    O = Operator
    cs, os_, ls = db.customer.schema, db.orders.schema, db.lineitem.schema
    c_scan = Scan(cs, db.customer, None, _b(_col(cs, "c_mktsegment"), O.Eq, Literal(ScalarValue.Utf8("BUILDING"))))
    o_scan = Scan(os_, db.orders, None, _b(_col(os_, "o_orderdate"), O.Lt, _date("1995-03-15")))
    l_scan = Scan(ls, db.lineitem, None, _b(_col(ls, "l_shipdate"), O.Gt, _date("1995-03-15")))
    j1 = HashJoinExec.try_new(c_scan, o_scan, JoinType.Inner, [(_col(cs, "c_custkey"), _col(os_, "o_custkey"))], None)
    j2 = HashJoinExec.try_new(j1, l_scan, JoinType.Inner, [(_col(j1.schema, "o_orderkey"), _col(ls, "l_orderkey"))], None)
    js = j2.schema
    rev = _b(_col(js, "l_extendedprice"), O.Mul, _b(_one20(), O.Sub, _col(js, "l_discount")))
    rt = pa.decimal128(38, 4)
    schema = pa.schema([("l_orderkey", pa.int64()), ("o_orderdate", pa.date32()), ("o_shippriority", pa.int64()), ("revenue", rt)])
    agg = HashAggregate(schema, j2, [_col(js, "l_orderkey"), _col(js, "o_orderdate"), _col(js, "o_shippriority")],
                        [SumAggregateExpr(rev, rt)])
    out = pa.schema([("l_orderkey", pa.int64()), ("revenue", rt), ("o_orderdate", pa.date32()), ("o_shippriority", pa.int64())])
    return Projection(out, agg, [Column("l_orderkey", 0), Column("revenue", 3), Column("o_orderdate", 1),
                                 Column("o_shippriority", 2)])


# ------------------------------------------------------------------------------------------------
# Q3 as a broadcast join over row-range shards (SURVEY 8e "Q3 joins"): what a distributed planner emits
# ------------------------------------------------------------------------------------------------
Q3_BUILD_SCHEMA = pa.schema([("o_orderkey", pa.int64()), ("o_orderdate", pa.date32()), ("o_shippriority", pa.int64())])


def q3_build_plan(db: Database) -> Projection:
    """J1 = customer(BUILDING) JOIN orders(o_orderdate < 1995-03-15) over THIS rank's shard of orders, projected to the
    three columns the second join and the aggregate need.  The ranks' results are all-gathered (broadcast build)."""
    O = Operator
    cs, os_ = db.customer.schema, db.orders.schema
    c_scan = Scan(cs, db.customer, None, _b(_col(cs, "c_mktsegment"), O.Eq, Literal(ScalarValue.Utf8("BUILDING"))))
    o_scan = Scan(os_, db.orders, None, _b(_col(os_, "o_orderdate"), O.Lt, _date("1995-03-15")))
    j1 = HashJoinExec.try_new(c_scan, o_scan, JoinType.Inner, [(_col(cs, "c_custkey"), _col(os_, "o_custkey"))], None)
    return Projection(Q3_BUILD_SCHEMA, j1, [_col(j1.schema, n) for n in Q3_BUILD_SCHEMA.names])


def q3_probe_plan(build, lineitem: MemoryTable) -> Projection:
    """J2 + aggregate of q3_plan over THIS rank's lineitem shard; same output schema.  build: the J1 rows of all ranks -- a
    plan producing them (Broadcast(q3_build_plan(...)): the whole step is then one native plan) or a MemoryTable holding
    them (the host-driven protocol)."""
    O = Operator
    ls = lineitem.schema
    b_scan = Scan(Q3_BUILD_SCHEMA, build, None, None) if isinstance(build, MemoryTable) else build
    l_scan = Scan(ls, lineitem, None, _b(_col(ls, "l_shipdate"), O.Gt, _date("1995-03-15")))
    j2 = HashJoinExec.try_new(b_scan, l_scan, JoinType.Inner, [(_col(Q3_BUILD_SCHEMA, "o_orderkey"), _col(ls, "l_orderkey"))], None)
    js = j2.schema
    rev = _b(_col(js, "l_extendedprice"), O.Mul, _b(_one20(), O.Sub, _col(js, "l_discount")))
    rt = pa.decimal128(38, 4)
    schema = pa.schema([("l_orderkey", pa.int64()), ("o_orderdate", pa.date32()), ("o_shippriority", pa.int64()), ("revenue", rt)])
    agg = HashAggregate(schema, j2, [_col(js, "l_orderkey"), _col(js, "o_orderdate"), _col(js, "o_shippriority")],
                        [SumAggregateExpr(rev, rt)])
    out = pa.schema([("l_orderkey", pa.int64()), ("revenue", rt), ("o_orderdate", pa.date32()), ("o_shippriority", pa.int64())])
    return Projection(out, agg, [Column("l_orderkey", 0), Column("revenue", 3), Column("o_orderdate", 1),
                                 Column("o_shippriority", 2)])


# ------------------------------------------------------------------------------------------------
# the complete Q1 / Q3 statements: ORDER BY (+ LIMIT) on top of the hot path (SURVEY 8f "next" #1)
# ------------------------------------------------------------------------------------------------
def q1_sorted_plan(db: Database):
    """q1.slt:1-22 with its `ORDER BY l_returnflag, l_linestatus` (planner: ascending, nulls_first = true)."""
    from .physical.plan import PhyscialSortExpr, Sort, SortOptions
    p = q1_plan(db)
    return Sort([PhyscialSortExpr(Column("l_returnflag", 0), SortOptions(False, True)),
                 PhyscialSortExpr(Column("l_linestatus", 1), SortOptions(False, True))], p)


def q3_top10_plan(db: Database):
    """q3.slt:1-24 with its `ORDER BY revenue DESC, o_orderdate LIMIT 10` (Sort::new_with_limit, planner/mod.rs:69-83)."""
    from .physical.plan import PhyscialSortExpr, Sort, SortOptions
    p = q3_plan(db)
    return Sort.new_with_limit([PhyscialSortExpr(Column("revenue", 1), SortOptions(True, True)),
                                PhyscialSortExpr(Column("o_orderdate", 2), SortOptions(False, True))], p, 10)
