#!/usr/bin/env python
"""bench.py -- TPC-H Q1 (default) / Q6 / Q3 rows/sec through the B200-native physical operators.

Contract (one JSON line on rank 0):
  value      input rows of the driving table (lineitem) / device time of one step, table resident in HBM
             (a step = one execution of the query's physical plan: Scan(pushed-down filter) ->
             [HashJoin ->] Aggregate -> Projection), whole job over all ranks, max over ranks.
  e2e        the same metric through the reference-facing operator API with HOST Arrow buffers:
             MemoryTable(host RecordBatches) -> H2D staging -> plan.execute() -> host RecordBatches.
  roofline   dominant kernel: algorithmic bytes of the layout the kernel reads / its CUDA-event duration
             (events recorded on the library's own stream by qgpu_profile_enable) vs MEASURED_PEAKS.json.
  cpu_baseline  oracle/qref_cpu.cpp (single-threaded C++ port of qurious's CPU steps) on a bounded sample.

`--impl reference` times that CPU port alone (the reference is Rust; no cargo/rustc in this image, so the
real binary cannot be built -- DESIGN.md "Oracle"); rank 0 only.

Workload at N GPUs: lineitem of TPC-H SF(sf*N) row-range sharded over the ranks (weak scaling, sf=10 per
GPU by default); partial aggregates are all-gathered over NCCL and merged identically on every rank.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--query", default="q1", choices=["q1", "q6", "q3", "groupby"])
    ap.add_argument("--rows", type=int, default=1_000_000_000, help="groupby workload: total rows (all GPUs)")
    ap.add_argument("--groups", type=int, default=100_000_000, help="groupby workload: distinct keys")
    ap.add_argument("--sf", type=float, default=10.0, help="TPC-H scale factor PER GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-rows", type=int, default=30_000_000)   # ~11 s of single-core CPU work for Q1
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--extra-queries", default="q6,q3,groupby",
                    help="comma list of further workloads whose device-resident leg (at the same N) is reported under 'queries'")
    ap.add_argument("--order-by", action="store_true", help="q1 / q3 with their ORDER BY (+ LIMIT 10) on the device (Sort, SURVEY 8f #1)")
    return ap.parse_args()


QUERY_COLUMNS = {
    "q1": {"lineitem": ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus",
                        "l_shipdate"]},
    "q6": {"lineitem": ["l_quantity", "l_extendedprice", "l_discount", "l_shipdate"]},
    "q3": {"customer": ["c_custkey", "c_mktsegment"],
           "orders": ["o_orderkey", "o_custkey", "o_orderdate", "o_shippriority"],
           "lineitem": ["l_orderkey", "l_extendedprice", "l_discount", "l_shipdate"]},
}
WORKLOAD = {"q1": "TPC-H Q1 group-by aggregate", "q6": "TPC-H Q6 filter + SUM",
            "q3": "TPC-H Q3 customer-orders-lineitem hash join + aggregate",
            "groupby": "synthetic int64-key group-by SUM/COUNT/MIN/MAX(v) AVG(f)"}
GROUPBY_SCHEMA = [("k", "int64"), ("v", "int64"), ("f", "float64")]


# ------------------------------------------------------------------------------------------------
# clocks (NVML) sampled DURING the timed regions
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    PERIOD_S = float(os.environ.get("QGPU_BENCH_SAMPLE_MS", "5")) / 1000.0
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, cuda_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thr = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(cuda_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.err = repr(e)
            self.nv = None

    def _run(self):
        # One NVML query pair every PERIOD_S while the timed region is open (the first ~1 ms after it opens).  Measured on
        # the host-bound Q3 step: 5 ms and 20 ms periods give the same step time (1.24 ms).
        nv = self.nv
        while not self._stop.is_set():
            if not self._active.wait(0.05):
                continue
            time.sleep(0.001)
            if self._active.is_set():
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    try:
                        mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:
                        mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(self.PERIOD_S)

    def region(self, on: bool):
        (self._active.set if on else self._active.clear)()

    def result(self):
        self._stop.set()
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def read_peak_live(torch):
    """Read-only streaming reference measured in this run: torch's int64 sum (a library reduction) over 1.68 GB in HBM.
    MEASURED_PEAKS.json's figure is a COPY (half reads, half writes); the scan kernels only read, and reads alone run
    faster than the copy -- a roofline.frac slightly above 1 against the copy peak is that difference, not an error."""
    try:
        x = torch.ones(210_000_000, dtype=torch.int64, device="cuda")
        for _ in range(3):
            x.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            x.sum()
        e1.record()
        torch.cuda.synchronize()
        gbs = x.numel() * 8 * 10 / (e0.elapsed_time(e1) / 1e3) / 1e9
        del x
        torch.cuda.empty_cache()
        return gbs
    except Exception:
        return None


PIN_FAILURES = []


def pin_batches(batches):
    """cudaHostRegister every Arrow buffer of the host batches so that H2D is a direct pinned DMA."""
    import torch
    rt = torch.cuda.cudart()
    regs = []
    seen = set()
    for b in batches:
        for col in b.columns:
            for buf in col.buffers():
                if buf is None or buf.size == 0 or buf.address in seen:
                    continue
                seen.add(buf.address)
                rc = int(rt.cudaHostRegister(buf.address, buf.size, 0))
                if rc == 0:
                    regs.append(buf.address)
                else:
                    PIN_FAILURES.append((buf.size, rc))
    return regs


def unpin(regs):
    import torch
    rt = torch.cuda.cudart()
    for a in regs:
        rt.cudaHostUnregister(a)


def batches_nbytes(batches):
    n = 0
    for b in batches:
        for col in b.columns:
            for buf in col.buffers():
                if buf is not None:
                    n += buf.size
    return n


def gen_groupby(rows, groups, device, row_range=None):
    """BASELINE.json configs[3]: k = bijectively scrambled uniform key over `groups` values, v in [-1e6, 1e6],
    f in [0, 1); counter-based like the TPC-H generator, so any row range can be generated independently."""
    import pyarrow as pa
    import torch
    from qurious_b200 import tpch
    lo, hi = row_range if row_range is not None else (0, rows)
    i = torch.arange(lo, hi, dtype=torch.int64, device=device)
    k = tpch._mix(tpch.rnd(50, i) % groups)
    v = tpch.uniform(51, i, -1_000_000, 1_000_000)
    f = (tpch.rnd(52, i) >> 9).to(torch.float64) / float(1 << 53)
    schema = pa.schema([("k", pa.int64()), ("v", pa.int64()), ("f", pa.float64())])
    return tpch.RawTable("t", schema, hi - lo, {"k": k, "v": v, "f": f}, {}, {})


def groupby_plan(table):
    import pyarrow as pa
    from qurious_b200.physical.expr import (AvgAggregateExpr, Column, CountAggregateExpr, MaxAggregateExpr, MinAggregateExpr,
                                            SumAggregateExpr)
    from qurious_b200.physical.plan import HashAggregate, Scan
    K, V, F = Column("k", 0), Column("v", 1), Column("f", 2)
    out = pa.schema([("k", pa.int64()), ("sum_v", pa.int64()), ("count_v", pa.int64()), ("min_v", pa.int64()),
                     ("max_v", pa.int64()), ("avg_f", pa.float64())])
    return HashAggregate(out, Scan(table.schema, table, None, None), [K],
                         [SumAggregateExpr(V, pa.int64()), CountAggregateExpr(V), MinAggregateExpr(V, pa.int64()),
                          MaxAggregateExpr(V, pa.int64()), AvgAggregateExpr(F, pa.float64(), pa.float64())])


ORDER_BY = False


def build_plan(query, tables):
    from qurious_b200 import tpch
    if query == "groupby":
        return groupby_plan(tables["t"])
    db = tpch.Database(0.0, tables.get("customer"), tables.get("orders"), tables.get("lineitem"))
    if ORDER_BY and query in ("q1", "q3"):
        return tpch.q1_sorted_plan(db) if query == "q1" else tpch.q3_top10_plan(db)
    return getattr(tpch, query + "_plan")(db)


def shard_range(total_rows, rank, world):
    return (total_rows * rank) // world, (total_rows * (rank + 1)) // world


def gen_raw(query, sf_total, device, rank, world):
    """RawTables for `query`; lineitem is this rank's row range, the (small) dimension tables are whole."""
    from qurious_b200 import tpch
    if query == "groupby":
        rows, groups = sf_total     # (total rows, groups) for this workload
        return {"t": gen_groupby(rows, groups, device, shard_range(rows, rank, world) if world > 1 else None)}
    cols = QUERY_COLUMNS[query]
    out = {}
    n_l = tpch.n_lineitems(sf_total)
    rr = shard_range(n_l, rank, world) if world > 1 else None
    out["lineitem"] = tpch.gen_lineitem(sf_total, device=device, columns=cols["lineitem"], row_range=rr)
    if "orders" in cols:
        orr = shard_range(tpch.n_orders(sf_total), rank, world) if world > 1 else None
        out["orders"] = tpch.gen_orders(sf_total, device=device, columns=cols["orders"], row_range=orr)
    if "customer" in cols:
        out["customer"] = tpch.gen_customer(sf_total, device=device, columns=cols["customer"])
    return out


def local_invariants(query, raw, torch, gather_keys=None):
    """Independent check values of THIS rank's shard, computed with plain torch int64 arithmetic over the generated
    columns (none of the library's code): they are summed over the ranks and compared with the merged result
    (`parity_check` in the JSON line).  -> {name: int}"""
    from qurious_b200 import tpch
    if query == "groupby":
        t = raw["t"].cols
        return {"rows": int(t["k"].numel()), "sum_v": int(t["v"].sum().item())}
    c = raw["lineitem"].cols
    if query == "q1":
        m = c["l_shipdate"] <= tpch.days("1998-09-02")
        return {"rows": int(m.sum().item()), "sum_qty": int(c["l_quantity"][m].sum().item()),
                "sum_base_price": int(c["l_extendedprice"][m].sum().item())}
    if query == "q6":
        m = ((c["l_shipdate"] >= tpch.days("1994-01-01")) & (c["l_shipdate"] < tpch.days("1995-01-01")) & (c["l_discount"] >= 5) &
             (c["l_discount"] <= 7) & (c["l_quantity"] < 2400))
        return {"revenue": int((c["l_extendedprice"][m] * c["l_discount"][m]).sum().item())}
    # q3: the qualifying orders of every rank's orders shard are needed by every rank (gather_keys: ragged all-gather)
    cust, o = raw["customer"], raw["orders"].cols
    d = tpch.days("1995-03-15")
    building = cust.cols["c_custkey"][cust.codes["c_mktsegment"] == cust.vocab["c_mktsegment"].index("BUILDING")]
    keys = o["o_orderkey"][(o["o_orderdate"] < d) & torch.isin(o["o_custkey"], building)]
    if gather_keys is not None:
        keys = gather_keys(keys)
    m = (c["l_shipdate"] > d) & torch.isin(c["l_orderkey"], keys)
    return {"revenue": int((c["l_extendedprice"][m] * (100 - c["l_discount"][m])).sum().item())}


def result_invariants(query, batches):
    """The same quantities read back from the query's (merged) result batches."""
    def raw_dec(v):
        return int(v.scaleb(-v.as_tuple().exponent))
    import pyarrow as pa
    if not batches:
        return {}
    t = pa.Table.from_batches(batches)
    if query == "groupby":
        return {"rows": sum(t["count_v"].to_pylist()), "sum_v": sum(t["sum_v"].to_pylist())}
    if query == "q1":
        return {"rows": sum(t["count_order"].to_pylist()), "sum_qty": sum(raw_dec(v) for v in t["sum_qty"].to_pylist()),
                "sum_base_price": sum(raw_dec(v) for v in t["sum_base_price"].to_pylist())}
    if query == "q6":
        v = t.column(0).to_pylist()[0]
        return {"revenue": raw_dec(v) if v is not None else 0}
    return {"revenue": sum(raw_dec(v) for v in t["revenue"].to_pylist())}


def metric_name(q):
    return "group-by rows/sec" if q == "groupby" else f"TPC-H {q.upper()} rows/sec"


def cpu_sample(query, rows_q1q6=2_000_000):
    """-> (callable running the CPU port once, driving-table rows it processes, description): a BOUNDED sample of the
    workload, generated by the same counter-based generators (so it is a prefix / a smaller scale factor of the same
    data), referenced columns only, 1024-row batches, one thread (the reference has no threads)."""
    from oracle import cpu_port
    from qurious_b200 import tpch
    if query in ("q1", "q6"):
        sf = rows_q1q6 / 6_001_215
        raw = tpch.gen_lineitem(sf, columns=QUERY_COLUMNS[query]["lineitem"])
        batches = tpch.to_arrow(raw, 1024 * 1024)
        return (lambda: getattr(cpu_port, query)(batches)), raw.rows, (
            f"first {raw.rows} rows of the same synthetic lineitem generator (SF{sf:.3f}), referenced columns only, "
            f"1024-row batches, single thread (the reference has no threads)")
    if query == "q3":
        sf = 0.5
        cols = QUERY_COLUMNS["q3"]
        t = {"customer": tpch.gen_customer(sf, columns=cols["customer"]), "orders": tpch.gen_orders(sf, columns=cols["orders"]),
             "lineitem": tpch.gen_lineitem(sf, columns=cols["lineitem"])}
        b = {k: tpch.to_arrow(v, 1024 * 1024) for k, v in t.items()}
        return (lambda: cpu_port.q3(b["customer"], b["orders"], b["lineitem"])), t["lineitem"].rows, (
            f"the same generators at SF{sf} ({t['lineitem'].rows} lineitem / {t['orders'].rows} orders / {t['customer'].rows} customer "
            f"rows), referenced columns only, 1024-row batches, single thread")
    rows, groups = 1_000_000, 100_000          # the workload's 10 rows per group
    raw = gen_groupby(rows, groups, "cpu")
    batches = tpch.to_arrow(raw, 1024 * 1024)
    return (lambda: cpu_port.groupby(batches)), rows, (
        f"{rows} rows / {groups} groups from the same generator (1/1000 of the workload, same rows per group), "
        f"1024-row batches, single thread")


def acero_context(query, batches):
    """SURVEY 8d "second line for context": the same query through Arrow C++ (pyarrow compute + Acero group_by) on ALL host
    cores.  It is NOT the reference (qurious is single-threaded arrow-rs) and not bit-identical to it: Arrow C++ refuses
    the reference's Decimal128(38, x) products (precision > 38), so the two wide Q1 products are formed in float64 -- the
    figure only says what a multi-threaded columnar CPU engine does with these rows."""
    import datetime
    import pyarrow as pa
    import pyarrow.compute as pc
    t = pa.Table.from_batches(batches)
    t0 = time.perf_counter()
    if query == "q6":
        from decimal import Decimal
        dec = pa.decimal128(15, 2)
        m = pc.and_(pc.and_(pc.greater_equal(t["l_shipdate"], datetime.date(1994, 1, 1)), pc.less(t["l_shipdate"], datetime.date(1995, 1, 1))),
                    pc.and_(pc.and_(pc.greater_equal(t["l_discount"], pa.scalar(Decimal("0.05"), dec)),
                                    pc.less_equal(t["l_discount"], pa.scalar(Decimal("0.07"), dec))),
                            pc.less(t["l_quantity"], pa.scalar(Decimal("24.00"), dec))))
        f = t.filter(m)
        out = pc.sum(pc.multiply(f["l_extendedprice"], f["l_discount"]))
        n_out = 1 if out.is_valid else 0
    else:
        f = t.filter(pc.less_equal(t["l_shipdate"], datetime.date(1998, 9, 2)))
        price, disc, tax = (pc.cast(f[c], pa.float64()) for c in ("l_extendedprice", "l_discount", "l_tax"))
        dp = pc.multiply(price, pc.subtract(1.0, disc))
        f = f.append_column("disc_price", dp).append_column("charge", pc.multiply(dp, pc.add(1.0, tax)))
        g = f.group_by(["l_returnflag", "l_linestatus"]).aggregate(
            [("l_quantity", "sum"), ("l_extendedprice", "sum"), ("disc_price", "sum"), ("charge", "sum"), ("l_quantity", "mean"),
             ("l_extendedprice", "mean"), ("l_discount", "mean"), ("l_quantity", "count")])
        n_out = g.num_rows
    dt = time.perf_counter() - t0
    return {"value": t.num_rows / dt, "unit": "rows/s", "cores": os.cpu_count(), "seconds": dt, "result_rows": n_out,
            "what": "Arrow C++ %s (pyarrow compute / Acero) on all host cores over the cpu_baseline sample; float64 for the "
                    "Decimal128(38,x) products Arrow C++ refuses; context only, not the reference" % pa.__version__}


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU port on host cores
# ------------------------------------------------------------------------------------------------
def workload_string(q, args, world, rows_total):
    """config.workload: the SAME string on both arms (the reference arm times a bounded sample of this workload per step)."""
    if q == "groupby":
        return f"{WORKLOAD[q]}: {rows_total} rows, {args.groups} groups, {world} GPU(s)"
    return (f"{WORKLOAD[q]} at SF{args.sf:g} per GPU ({rows_total} lineitem rows total, "
            f"row-range sharded over {world} GPU(s))")


def run_reference(args):
    """The CPU arm: oracle/qref_cpu.cpp (kind "port": the reference is Rust, no cargo / rustc here) on the host cores, one
    thread like the reference, each step the SAME bounded sample `cpu_baseline` of the B200 arm uses (--cpu-sample-rows,
    30 M lineitem rows for Q1 / Q6: ~12 s per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from qurious_b200 import tpch
    world = max(1, args.gpus)
    rows_total = args.rows if args.query == "groupby" else tpch.n_lineitems(args.sf * world)
    fn, n_rows, sample = cpu_sample(args.query, min(args.cpu_sample_rows, rows_total))
    for _ in range(1 if args.warmup > 0 else 0):      # one warm-up pass (each is ~12 s of single-core work)
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    rows_s = n_rows * args.steps / dt
    line = {"impl": "reference", "metric": metric_name(args.query), "value": rows_s, "unit": "rows/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong" if args.query == "groupby" else "weak", "vs_baseline": None,
            "dtype": "i64/i128 (exact decimal)", "data": "synthetic",
            "config": {"workload": workload_string(args.query, args, world, rows_total)},
            "cpu_baseline": {"value": rows_s, "unit": "rows/s", "cores": 1, "kind": "port", "sample": sample,
                             "sample_rows_per_step": n_rows, "warmup_passes_run": 1 if args.warmup > 0 else 0},
            "e2e": {"value": rows_s, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def kernel_bytes(kernel, q, per_table, driving, rows_local, groups_local):
    """Algorithmic bytes ONE launch of `kernel` moves, or None when the kernel does not stream a table (no
    roofline fraction is quoted for it): the fused scan kernels read every referenced column of the table they scan once
    (FM_EMIT = J1's probe scan over orders, everything else the driving table); the radix passes move 3 x 8 B tuples."""
    if kernel.startswith("k_fused_scan_agg"):       # (_spec: ahead-of-time shape, _jit: run-time shape, <FM_*>: generic body)
        return per_table.get("orders" if "FM_EMIT" in kernel else driving)
    if q == "groupby":
        # tuples are 3 x 8 B; groups leave as 6 x 8 B
        # (kernel names carry template arguments and the _tma suffix of the pipelined variants: match on the stem)
        for stem, b in (("k_radix_agg", 24 * rows_local + 48 * groups_local), ("k_radix_scatter", 48 * rows_local),
                        ("k_radix_hist1", 8 * rows_local), ("k_radix_hist2", 8 * rows_local)):
            if kernel.startswith(stem):
                return b
        return None
    return None


def algorithmic_bytes(query, dev_tables):
    """Bytes of every referenced column in the layout the kernels read (after ingest narrowing), once."""
    total = 0
    per_table = {}
    for name, mt in dev_tables.items():
        dev = mt._dev
        b = sum(dev.column_bytes(i) for i in range(len(dev.schema)))
        per_table[name] = b
        total += b
    return total, per_table


class Pipelined:
    """Depth-2 pipeline of asynchronous executions (qgpu_plan_execute_device_async / _sharded_device async): step i is
    queued, then the result of step i-1 is waited for (its row count is read: DeviceTable.wait) and released -- the host
    consumes every result while the GPU already runs the next step.  Plans whose last operator is not the fused dense
    aggregate resolve inside the call, i.e. run synchronously as before."""

    def __init__(self, launch):
        self.launch, self.inflight, self.rows = launch, [], None

    def __call__(self):
        self.inflight.append(self.launch())
        if len(self.inflight) > 1:
            self._retire()

    def _retire(self):
        t = self.inflight.pop(0)
        t.wait()
        self.rows = t.num_rows
        t.free()

    def drain(self):
        while self.inflight:
            self._retire()


def run_query_device(ctx, step, steps, warmup, sampler, torch, stream):
    """K timed steps, table resident in HBM; returns (ms_total, per-kernel profile, launches per step)."""
    drain = getattr(step, "drain", lambda: None)
    for _ in range(warmup):
        step()
    drain()
    # event pairs only around launches of >= 64 blocks (the scan / probe / radix kernels): bracketing every single-block
    # helper as well costs the Q1 SF10 step 7 % (0.463 vs 0.431 ms, same box; QGPU_BENCH_PROFILE_ALL=1 restores that)
    ctx.profile(True, 0 if os.environ.get("QGPU_BENCH_PROFILE_ALL") else 64)
    ctx.profile_report()
    l0 = ctx.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    sampler.region(True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(min(steps, 64))]   # per-step marks (first 64 steps)
    e0.record(stream)
    for i in range(steps):
        step()
        if i < len(marks):
            marks[i].record(stream)
    drain()                      # the last results are awaited inside the timed region: their epilogues run on the library's side stream
    e1.record(stream)
    torch.cuda.synchronize()
    sampler.region(False)
    barrier()
    ms = e0.elapsed_time(e1)
    run_query_device.step_ms = [round(a.elapsed_time(b), 4) for a, b in zip([e0] + marks[:-1], marks)]
    prof = ctx.profile_report()
    ctx.profile(False)
    launches = (ctx.kernel_launches() - l0) / max(steps, 1)
    return ms, prof, launches


_dist = None


def barrier():
    if _dist is not None:
        _dist.barrier()


def _gather_ragged_i64(t, world, torch):
    """all ranks' int64 vectors concatenated (NCCL all-gather of the sizes, then of the padded vectors)"""
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    ns = torch.empty(world, dtype=torch.int64, device=t.device)
    _dist.all_gather_into_tensor(ns, n)
    sizes = [int(x) for x in ns.tolist()]
    pad = torch.zeros(max(sizes + [1]), dtype=torch.int64, device=t.device)
    pad[:t.numel()] = t
    out = torch.empty(world * pad.numel(), dtype=torch.int64, device=t.device)
    _dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * pad.numel(): r * pad.numel() + sizes[r]] for r in range(world)])


def parity_check(q, expect_local, result_of, world, torch):
    """One more execution of the step's plan; its (merged) result is compared with invariants computed independently
    by torch over every rank's generated columns (summed over the ranks with all_reduce): sum of the group counts =
    qualifying rows, sum of the group sums = the plain column sums."""
    try:
        names = sorted(expect_local)
        exp = torch.tensor([expect_local[k] for k in names], dtype=torch.int64, device="cuda")
        if world > 1:
            _dist.all_reduce(exp)
        res = result_of()
        if isinstance(res, list):
            got = result_invariants(q, res)
            got_t = torch.tensor([got.get(k, 0) for k in names], dtype=torch.int64, device="cuda")
            if world > 1 and q == "q3":     # the distributed Q3 result stays sharded: every group on exactly one rank
                _dist.all_reduce(got_t)
        else:                               # group-by: the result stays in HBM (10^8 groups); this rank's share of the groups
            from qurious_b200.distributed import column_bytes_tensor
            cnt = column_bytes_tensor(res, 2)[0].view(torch.int64)
            sv = column_bytes_tensor(res, 1)[0].view(torch.int64)
            got_t = torch.stack([cnt.sum(), sv.sum()])           # names: rows, sum_v
            res.free()
            if world > 1:
                _dist.all_reduce(got_t)
        ok = bool(torch.equal(exp, got_t))
        return {"ok": ok, "checked": names, "expected": [int(x) for x in exp.tolist()], "got": [int(x) for x in got_t.tolist()],
                "how": "torch int64 sums over the generated columns of every shard (all_reduce) vs the query's merged result"}
    except Exception as e:                  # a failing check must be visible, not fatal for the measurement
        return {"ok": False, "error": repr(e)[:300]}


def setup_leg(q, args, ctx, rank, world, torch):
    """Tables of `q` resident in HBM (this rank's shard at N > 1), its plan, the step callable of the device-resident leg
    and the independent check values.  -> dict"""
    from qurious_b200 import tpch
    sf_total = args.sf * world if q != "groupby" else (args.rows, args.groups)
    driving = "t" if q == "groupby" else "lineitem"
    raw = gen_raw(q, sf_total, "cuda", rank, world)
    rows_local = raw[driving].rows
    expect = local_invariants(q, raw, torch, (lambda k: _gather_ragged_i64(k, world, torch)) if world > 1 else None)
    dev_tables = {k: tpch.to_device_table(ctx, v) for k, v in raw.items()}
    del raw
    torch.cuda.empty_cache()        # the generator's temporaries go back to the driver: HBM is for the tables
    plan = build_plan(q, dev_tables)
    leg = {"q": q, "plan": plan, "tables": dev_tables, "rows_local": rows_local, "expect": expect, "driving": driving,
           "sf_total": sf_total, "lo": 0, "strategy": lambda: plan.last_strategy()}
    if world > 1:
        from qurious_b200 import distributed as qd
        lo = 0 if q == "groupby" else shard_range(tpch.n_lineitems(sf_total), rank, world)[0]
        leg["lo"] = lo
        if q == "groupby":
            # config 4: one kernel partitions the rows by key hash and stores every tuple into its owner's HBM over
            # NVLink (peer-to-peer), then a purely local aggregate (groups are rank-disjoint); result = concatenation
            xg = qd.ExchangeGroupBy(ctx, plan, world, rank)

            def step():
                xg.execute_device().free()
            leg["strategy"] = lambda: "%s x%d -> %s" % (xg.last_path, world, plan.last_strategy())
            leg["result_of"] = xg.execute_device
        elif q == "q3":
            # orders and lineitem row-range sharded, customer replicated; ONE native plan per rank (csrc/exchange.cu):
            # FinalAggregate <- Aggregate <- Join(Broadcast(J1 over the orders shard), lineitem shard): the J1 rows of all
            # ranks reach every GPU in one grouped NCCL exchange, groups straddling a shard boundary are merged by a hash
            # exchange of the partial groups; the result stays sharded
            # Broadcast with key-range pruning: a J1 row only travels to the ranks whose lineitem shard can hold its order key
            bj = qd.BroadcastJoinAggregate(ctx, tpch.q3_build_plan(tpch.Database(0.0, dev_tables["customer"], dev_tables["orders"], None)),
                                           lambda b: tpch.q3_probe_plan(b, dev_tables["lineitem"]), world,
                                           prune=(0, dev_tables["lineitem"], dev_tables["lineitem"].schema.get_field_index("l_orderkey")))

            host_t = {"exec": 0.0, "free": 0.0, "n": 0}

            def step():
                t0 = time.perf_counter()
                t = bj.execute_device()
                t1 = time.perf_counter()
                t.free()
                host_t["exec"] += t1 - t0
                host_t["free"] += time.perf_counter() - t1
                host_t["n"] += 1
                if os.environ.get("QGPU_BENCH_HOSTTIME") and host_t["n"] % 10 == 0:
                    print("[bench] q3 host ms per step: execute_device %.3f, free %.3f" % (1e3 * host_t["exec"] / host_t["n"], 1e3 * host_t["free"] / host_t["n"]), file=sys.stderr)
            leg["strategy"] = lambda: bj.last_strategy
            leg["result_of"] = bj.execute
        else:
            # Q1 / Q6: scan kernel + ONE epilogue kernel per step; the epilogue stores the state block into every peer's
            # buffer over NVLink, waits on the peers' flags, merges and finalises (no collective call, no host round trip)
            sharded = qd.ShardedAggregate(ctx, plan, lo, world)
            sharded.execute_device().free()            # records the strategy
            step = Pipelined(lambda: sharded.execute_device(wait=False))
            leg["result_of"] = sharded.execute
    else:
        plan.execute_device(ctx).free()                # records the strategy
        step = Pipelined(lambda: plan.execute_device_async(ctx))
        leg["result_of"] = (lambda: plan.execute_device(ctx)) if q == "groupby" else (lambda: plan.execute(ctx))
    leg["step"] = step
    return leg


def free_leg(leg, ctx, torch):
    leg["plan"].release()
    for t in leg["tables"].values():
        if t._dev is not None:
            t._dev.free()
    leg.clear()
    ctx.release_cached_memory()
    torch.cuda.empty_cache()


def run_b200(args):
    global _dist
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the ONE JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        _dist = dist
    from qurious_b200 import _lib, tpch
    ctx = _lib.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank)
    peak, peak_src = measured_peak_gbs()
    q = args.query
    if world > 1:
        from qurious_b200 import distributed as qd
        qd.init_comm(ctx)           # the library-owned communicator (NCCL + symmetric peer buffers)

    # ---- device-resident leg -----------------------------------------------------------------
    leg = setup_leg(q, args, ctx, rank, world, torch)
    sf_total, driving, rows_local, dev_tables, plan, lo = (leg[k] for k in ("sf_total", "driving", "rows_local", "tables", "plan", "lo"))
    step, expect, result_of = leg["step"], leg["expect"], leg["result_of"]
    ms, prof, launches = run_query_device(ctx, step, args.steps, args.warmup, sampler, torch, stream)
    parity = parity_check(q, expect, result_of, world, torch)
    main_step_ms = list(getattr(run_query_device, "step_ms", []))[:64]     # (the extra-query legs below overwrite the attribute)
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    rows_t = torch.tensor([rows_local], dtype=torch.int64, device="cuda")
    if world > 1:
        _dist.all_reduce(t_ms, op=_dist.ReduceOp.MAX)
        _dist.all_reduce(rows_t, op=_dist.ReduceOp.SUM)
    ms_max, rows_total = float(t_ms.item()), int(rows_t.item())
    value = rows_total * args.steps / (ms_max / 1e3)
    strategy = leg["strategy"]()
    alg_bytes, per_table = algorithmic_bytes(q, dev_tables)
    # dominant kernel by total device time
    prof_sorted = sorted(prof, key=lambda r: -r[2])
    top = prof_sorted[0] if prof_sorted else ("none", 0, 0.0, 0.0)
    top_ms = top[2] / max(top[1], 1)
    top_share = top[2] / max(ms, 1e-9)                 # of the timed region's device time
    top_bytes = kernel_bytes(top[0], q, per_table, driving, rows_local, args.groups // world)
    achieved = top_bytes / (top_ms / 1e3) / 1e9 if (top_ms > 0 and top_bytes) else None
    traffic, traffic_src = None, None      # measured DRAM bytes per launch of that kernel at this size (ncu), when captured
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        hit = tj.get(f"{top[0]}|{q}|{rows_local}")
        if hit:
            traffic, traffic_src = hit["bytes"], hit["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved is not None else None, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": top_bytes, "bytes_per_row": (top_bytes / max(rows_local, 1)) if top_bytes else None,
                "kernel_ms_avg": top_ms, "kernel_share_of_step": top_share, "launches_of_kernel_per_step": top[1] / args.steps,
                "step_frac_of_roofline": (alg_bytes / (ms / args.steps / 1e3) / 1e9) / peak}
    ctx.release_cached_memory()
    rp = read_peak_live(torch)
    if rp and achieved is not None:
        roofline["read_only_reference"] = {"gbs": rp, "frac": achieved / rp,
                                           "how": "torch int64 sum over 1.68 GB, measured in this run (reads only; the peak above is a copy)"}

    # ---- the same plan through the other two tile bodies -------------------------------------------------------------------
    # The headline runs the AHEAD-OF-TIME specialised shape (fused.cu: SIG_Q1 / SIG_Q6).  Every other DENSE plan gets the
    # same SpecBody instantiated for its shape at RUN TIME (NVRTC, fused_jit.cu): roofline_jit forces that path for this very
    # workload (QGPU_FUSED_NOAOT=1).  roofline_generic is the interpreted body that remains the fallback (QGPU_FUSED_GENERIC=1:
    # no NVRTC on the host, a failed compilation) and that the HASH / join-probe modes still run.
    roofline_generic = roofline_jit = None
    if world == 1 and q in ("q1", "q6") and not os.environ.get("QGPU_BENCH_SKIP_GENERIC"):
        def variant(env_name, how):
            try:
                os.environ[env_name] = "1"
                gplan = build_plan(q, dev_tables)           # a fresh plan object: the analysis (and the kernel choice) is per plan
                gplan.execute_device(ctx).free()
                gstep = Pipelined(lambda: gplan.execute_device_async(ctx))
                gsteps = max(3, min(args.steps, 20))
                gms, gprof, _ = run_query_device(ctx, gstep, gsteps, min(args.warmup, 3), sampler, torch, stream)
                gtop = sorted(gprof, key=lambda r: -r[2])[0]
                gtop_ms = gtop[2] / max(gtop[1], 1)
                gbytes = kernel_bytes(gtop[0], q, per_table, driving, rows_local, 0)
                gach = gbytes / (gtop_ms / 1e3) / 1e9 if gbytes and gtop_ms > 0 else None
                out = {"kernel": gtop[0], "kernel_ms_avg": gtop_ms, "achieved": gach, "peak": peak, "unit": "GB/s",
                       "frac": (gach / peak) if gach else None, "ms_per_step": gms / gsteps, "strategy": gplan.last_strategy(), "how": how}
                gplan.release()
                return out
            except Exception as e:      # context figure: never fatal
                return {"error": repr(e)[:300]}
            finally:
                os.environ.pop(env_name, None)
        roofline_jit = variant("QGPU_FUSED_NOAOT", "the same plan and tables with QGPU_FUSED_NOAOT=1: SpecBody instantiated for this shape by NVRTC at run time")
        roofline_generic = variant("QGPU_FUSED_GENERIC", "the same plan and tables with QGPU_FUSED_GENERIC=1: the interpreted tile body")

    # ---- end-to-end leg: host Arrow buffers -> operators -> host RecordBatches ---------------------
    e2e = None
    cpu = None
    host_batches = None
    ctx.release_cached_memory()     # the device leg's re-usable blocks: the generator below needs the HBM
    if not args.no_e2e and q != "groupby":
        raw_h = gen_raw(q, sf_total, "cuda", rank, world)
        host = {}
        for k, v in raw_h.items():
            for d in (v.cols, v.codes):
                for c in list(d):
                    d[c] = d[c].cpu()
            host[k] = tpch.to_arrow(v, None)
        del raw_h
        torch.cuda.empty_cache()
        host_batches = host
        h2d = sum(batches_nbytes(b) for b in host.values())
        from qurious_b200.physical.plan import MemoryTable

        def measure_e2e(tables, steps, label, path):
            """`steps` timed end-to-end steps over the host tables `tables` ({name: [RecordBatch]})."""
            phase = {"upload_ms": 0.0, "execute_ms": 0.0, "n": 0}

            def one_e2e():
                t_a = time.perf_counter()
                tabs = {k: MemoryTable.try_new(b[0].schema, b) for k, b in tables.items()}
                for t in tabs.values():            # append (retain) every batch, then ONE staged upload of the table (ingest.cu)
                    t.device_table(ctx).flush()
                stream.synchronize()
                t_b = time.perf_counter()
                p = build_plan(q, tabs)
                if world > 1 and q == "q3":
                    out = qd.BroadcastJoinAggregate(ctx, tpch.q3_build_plan(tpch.Database(0.0, tabs["customer"], tabs["orders"], None)),
                                                    lambda b: tpch.q3_probe_plan(b, tabs["lineitem"]), world,
                                                    prune=(0, tabs["lineitem"], tabs["lineitem"].schema.get_field_index("l_orderkey"))).execute()
                elif world > 1:
                    out = qd.ShardedAggregate(ctx, p, lo, world).execute()
                else:
                    out = p.execute(ctx)
                d2h = batches_nbytes(out)
                t_c = time.perf_counter()
                phase["upload_ms"] += 1e3 * (t_b - t_a)
                phase["execute_ms"] += 1e3 * (t_c - t_b)
                phase["n"] += 1
                p.release()
                for t in tabs.values():
                    if t._dev is not None:
                        t._dev.free()
                return d2h
            for _ in range(min(args.warmup, 2)):
                d2h = one_e2e()
            phase.update(upload_ms=0.0, execute_ms=0.0, n=0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            torch.cuda.synchronize()
            sampler.region(True)
            t0 = time.perf_counter()
            e0.record(stream)
            for _ in range(steps):
                d2h = one_e2e()
            e1.record(stream)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            sampler.region(False)
            barrier()
            ems = max(e0.elapsed_time(e1), wall * 1e3)   # host-side staging is part of the step
            t_e = torch.tensor([ems], dtype=torch.float64, device="cuda")
            if world > 1:
                _dist.all_reduce(t_e, op=_dist.ReduceOp.MAX)
            return {"label": label, "value": rows_total * steps / (float(t_e.item()) / 1e3), "unit": "rows/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": steps,
                    "ms_per_step": float(t_e.item()) / steps,
                    "upload_ms_per_step": phase["upload_ms"] / max(phase["n"], 1),
                    "execute_ms_per_step": phase["execute_ms"] / max(phase["n"], 1),
                    "batches_per_step": sum(len(b) for b in tables.values()), "host_threads": os.cpu_count(), "path": path}

        staged = ("qgpu_table_append (batches retained) -> one staged upload per table: host worker threads gather the batches "
                  "through a pinned ring, Decimal128(15,2) narrowed to int64 on the way (46 instead of 78 B/row cross PCIe for Q1) "
                  "-> plan.execute() -> host RecordBatches")
        paths = {}
        # (i) the call a user of the reference makes: pageable Arrow buffers, one RecordBatch per table
        paths["pageable_single_batch"] = measure_e2e(host, args.e2e_steps, "pageable_single_batch",
                                                     "MemoryTable(host RecordBatches, pageable, 1 batch per table) -> " + staged)
        # (iii) the reference's own granularity: 1024-row batches (datasource/file/csv.rs:34-72), pageable
        if world == 1 and not os.environ.get("QGPU_BENCH_SKIP_1024"):
            small = {k: [b[0].slice(o, 1024) for o in range(0, b[0].num_rows, 1024)] for k, b in host.items()}
            r = measure_e2e(small, max(1, min(args.e2e_steps, 2)), "pageable_1024_row_batches",
                            "MemoryTable(host RecordBatches, pageable, 1024-row batches) -> qgpu_table_append_stream -> " + staged)
            # what pyarrow alone needs to hand these batches over the Arrow C stream interface (export + release, no upload):
            # the floor of this path in this Python harness, outside the library
            from pyarrow.cffi import ffi as _ffi
            import pyarrow as pa
            t0 = time.perf_counter()
            for k, bl in small.items():
                cs = _ffi.new("struct ArrowArrayStream*")
                pa.RecordBatchReader.from_batches(bl[0].schema, bl)._export_to_c(int(_ffi.cast("uintptr_t", cs)))
                arr = _ffi.new("struct ArrowArray*")
                while True:
                    cs.get_next(cs, arr)
                    if arr.release == _ffi.NULL:
                        break
                    arr.release(arr)
                cs.release(cs)
            r["arrow_c_stream_export_floor_ms"] = 1e3 * (time.perf_counter() - t0)
            paths["pageable_1024_row_batches"] = r
            del small
        # (iv) file -> HBM: a '|'-delimited text file of the first rows of this run's lineitem (written outside the timed region by
        # Arrow C++'s CSV writer), read and PARSED ON THE GPU (qgpu_table_append_csv_file, csrc/csv.cu) -- the reference's
        # `COPY lineitem FROM 'lineitem.tbl' (DELIMITER '|')` route without any host-side Arrow batches
        if world == 1 and q in ("q1", "q6") and not os.environ.get("QGPU_BENCH_SKIP_CSV"):
            try:
                import tempfile
                import pyarrow.csv as pacsv
                from qurious_b200 import _lib as qlib
                n_s = min(6_001_215, host["lineitem"][0].num_rows)
                sample = host["lineitem"][0].slice(0, n_s)
                tmpdir = tempfile.mkdtemp(prefix="qgpu_bench_")
                path = os.path.join(tmpdir, "lineitem.tbl")
                pacsv.write_csv(sample, path, pacsv.WriteOptions(include_header=False, delimiter="|", quoting_style="none"))
                fbytes = os.path.getsize(path)

                def one_file():
                    dev = qlib.DeviceTable.create(ctx, sample.schema)
                    dev.append_csv(path, has_header=False, delimiter="|")
                    mt = MemoryTable.from_device_table(dev)
                    p = build_plan(q, {"lineitem": mt})
                    out = p.execute(ctx)
                    p.release()
                    dev.free()
                    return out
                one_file()
                t0 = time.perf_counter()
                reps = 3
                for _ in range(reps):
                    one_file()
                dt = (time.perf_counter() - t0) / reps
                paths["tbl_file_sample"] = {"label": "tbl_file_sample", "value": n_s / dt, "unit": "rows/s", "rows": n_s, "file_bytes": fbytes,
                                            "ms_per_step": dt * 1e3, "text_gb_per_s": fbytes / dt / 1e9, "steps": reps,
                                            "path": "lineitem.tbl (first %d rows, '|'-delimited text, page cache) -> qgpu_table_append_csv_file: pread into "
                                                    "the pinned ring -> H2D -> split + parse kernels -> resident columns -> plan.execute() -> host "
                                                    "RecordBatches" % n_s}
                os.remove(path)
                os.rmdir(tmpdir)
            except Exception as e:      # context figure: never fatal
                paths["tbl_file_sample"] = {"error": repr(e)[:300]}
        # (ii) page-locked sources (cudaHostRegister outside the timed region): untransformed columns are DMA'd directly
        regs = []
        for k in host:
            regs += pin_batches(host[k])
        r = measure_e2e(host, args.e2e_steps, "pinned_single_batch",
                        "MemoryTable(host RecordBatches, page-locked, 1 batch per table) -> " + staged)
        r["host_buffers_pinned"], r["host_buffers_pin_failed"] = len(regs), len(PIN_FAILURES)
        paths["pinned_single_batch"] = r
        unpin(regs)
        # headline: page-locked host buffers, as the bench contract words it ("the host->device copy of that step's inputs from
        # pinned host memory"); the pageable and 1024-row-batch figures stand beside it under e2e.paths
        e2e = dict(paths["pinned_single_batch"])
        e2e["paths"] = {k: v for k, v in paths.items() if k != "pinned_single_batch"}

    # ---- CPU baseline (rank 0, N=1): the C++ port on a bounded sample of the same workload -----------
    if not args.no_cpu and rank == 0 and world == 1:
        if q in ("q1", "q6") and host_batches is not None:
            from oracle import cpu_port
            n = min(args.cpu_sample_rows, host_batches["lineitem"][0].num_rows)
            sample = [host_batches["lineitem"][0].slice(0, n)]
            t0 = time.perf_counter()
            getattr(cpu_port, q)(sample)
            dt = time.perf_counter() - t0
            cpu = {"value": n / dt, "unit": "rows/s", "cores": 1, "kind": "port",
                   "sample": f"first {n} rows of this run's lineitem (referenced columns only), 1024-row batches, "
                             f"{dt:.1f} s on 1 of {os.cpu_count()} host cores (the reference is single-threaded)"}
            try:
                cpu["context_arrow_cpp_all_cores"] = acero_context(q, sample)
            except Exception as e:      # context only: never fatal
                cpu["context_arrow_cpp_all_cores"] = {"error": repr(e)[:200]}
        else:
            fn, n, desc = cpu_sample(q)
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            cpu = {"value": n / dt, "unit": "rows/s", "cores": 1, "kind": "port",
                   "sample": f"{desc}; {dt:.1f} s on 1 of {os.cpu_count()} host cores"}

    # ---- the other workloads of BASELINE.json (Q6, Q3, the 1 B-row group-by), device-resident leg, at the same N: "queries" ----
    extra = {}
    if args.extra_queries:
        free_leg(leg, ctx, torch)       # the main workload's tables: HBM for the next ones
        for xq in [x.strip() for x in args.extra_queries.split(",")]:
            if not xq or xq == q or xq not in WORKLOAD:
                continue
            try:
                xleg = setup_leg(xq, args, ctx, rank, world, torch)
                xsteps = max(1, min(args.steps, 5 if xq == "groupby" else 20))
                xwarm = min(args.warmup, 3) if xq == "groupby" else args.warmup
                xms, xprof, xlaunches = run_query_device(ctx, xleg["step"], xsteps, xwarm, sampler, torch, stream)
                xstep_ms = list(getattr(run_query_device, "step_ms", []))[:20]
                xparity = parity_check(xq, xleg["expect"], xleg["result_of"], world, torch)
                xt = torch.tensor([xms], dtype=torch.float64, device="cuda")
                xr = torch.tensor([xleg["rows_local"]], dtype=torch.int64, device="cuda")
                if world > 1:
                    _dist.all_reduce(xt, op=_dist.ReduceOp.MAX)
                    _dist.all_reduce(xr, op=_dist.ReduceOp.SUM)
                xms_max, xrows = float(xt.item()), int(xr.item())
                xsorted = sorted(xprof, key=lambda r: -r[2])
                xtop = xsorted[0] if xsorted else ("none", 0, 0.0, 0.0)
                xtop_ms = xtop[2] / max(xtop[1], 1)
                _, xper = algorithmic_bytes(xq, xleg["tables"])
                xbytes = kernel_bytes(xtop[0], xq, xper, xleg["driving"], xleg["rows_local"], args.groups // world)
                xach = xbytes / (xtop_ms / 1e3) / 1e9 if (xtop_ms > 0 and xbytes) else None
                extra[xq] = {"metric": metric_name(xq), "value": xrows * xsteps / (xms_max / 1e3), "unit": "rows/s",
                             "ms_per_step": xms_max / xsteps, "steps": xsteps, "rows": xrows,
                             "scaling": "strong" if xq == "groupby" else "weak", "strategy": xleg["strategy"](),
                             "roofline": {"bound": "hbm", "kernel": xtop[0], "kernel_ms_avg": xtop_ms, "achieved": xach, "peak": peak,
                                          "unit": "GB/s", "frac": (xach / peak) if xach is not None else None,
                                          "algorithmic_bytes_per_launch": xbytes, "kernel_share_of_step": xtop[2] / max(xms, 1e-9)},
                             "gpu_launches_per_step": xlaunches, "step_ms": xstep_ms, "parity_check": xparity,
                             "kernels": [{"name": r[0], "launches": r[1], "total_ms": round(r[2], 4)} for r in xsorted[:6]]}
                free_leg(xleg, ctx, torch)
            except Exception as e:      # never let an extra workload take the headline line down
                extra[xq] = {"error": repr(e)[:300]}
    clocks = sampler.result()
    if rank == 0:
        line = {"metric": metric_name(q), "value": value, "unit": "rows/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "strong" if q == "groupby" else "weak",   # the group-by's row count is fixed, TPC-H is SF per GPU
                "vs_baseline": None, "dtype": "i64/i128 (exact decimal)", "data": "synthetic",
                "config": {"workload": workload_string(q, args, world, rows_total),
                           "l2": "inputs larger than L2 (resident columns >> 126 MB), no flush needed",
                           "strategy": strategy, "rows_per_gpu": rows_local},
                "roofline": roofline, "roofline_jit": roofline_jit, "roofline_generic": roofline_generic, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(round(launches * args.steps)),
                "gpu_launches_per_step": launches, "clocks": clocks,
                "step_ms": main_step_ms, "parity_check": parity,
                "kernels": [{"name": r[0], "launches": r[1], "total_ms": round(r[2], 4)} for r in prof_sorted[:8]]}
        if extra:
            line["queries"] = extra
        print(json.dumps(line))
    if world > 1:
        _dist.destroy_process_group()


def main():
    global ORDER_BY
    args = parse_args()
    ORDER_BY = args.order_by
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
