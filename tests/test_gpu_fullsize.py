"""Parity at BASELINE.json's FULL sizes (TPC-H SF10 Q1 / Q6 / Q3 on one B200, 1 B rows / 100 M groups), where the oracle
cannot run in seconds: the same seeded generator produces the columns in HBM, and plain torch int64 arithmetic over them
(exact: every sum stays below 2^63 at these sizes) is the independent checker -- whole-result equality for Q6 and Q1,
and EXACT PER-GROUP equality for Q3 and the 1 B-row group-by (torch.unique(return_inverse) + index_add_ /
scatter_reduce as the independent per-group aggregator).  The expected values follow the reference's arithmetic: decimal mul = raw product, scale s1+s2
(binary.rs:51-68 / arrow mul_wrapping), 1 - d = 100 - raw(d) at scale 2, decimal AVG = sum * 10^4 / count truncated
(avg.rs:89-116), COUNT = rows (count.rs:40-48)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bench  # noqa: E402
from qurious_b200 import tpch  # noqa: E402
from qurious_b200.distributed import column_bytes_tensor  # noqa: E402
from tests.cases import rows_of  # noqa: E402

pytestmark = pytest.mark.gpu
SF = 10.0


def raw_dec(v):
    return int(v.scaleb(-v.as_tuple().exponent))


def trunc_div(a: int, b: int) -> int:
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


@pytest.fixture(scope="module")
def lineitem():
    cols = sorted(set(bench.QUERY_COLUMNS["q1"]["lineitem"]) | set(bench.QUERY_COLUMNS["q3"]["lineitem"]))
    t = tpch.gen_lineitem(SF, device="cuda", columns=cols)
    assert t.rows == 59_986_052
    yield t
    del t
    torch.cuda.empty_cache()


def test_q6_sf10_whole_result(gpu_ctx, lineitem):
    c = lineitem.cols
    m = ((c["l_shipdate"] >= tpch.days("1994-01-01")) & (c["l_shipdate"] < tpch.days("1995-01-01")) & (c["l_discount"] >= 5) &
         (c["l_discount"] <= 7) & (c["l_quantity"] < 2400))
    expect = int((c["l_extendedprice"][m] * c["l_discount"][m]).sum().item())
    mt = tpch.to_device_table(gpu_ctx, lineitem)
    plan = tpch.q6_plan(tpch.Database(SF, None, None, mt))
    got = rows_of(plan.execute(gpu_ctx))
    assert "fused_scan_agg" in plan.last_strategy()
    assert len(got) == 1 and raw_dec(got[0][0]) == expect
    mt._dev.free()


def test_q1_sf10_whole_result(gpu_ctx, lineitem):
    c, codes, vocab = lineitem.cols, lineitem.codes, lineitem.vocab
    keep = c["l_shipdate"] <= tpch.days("1998-09-02")
    expect = {}
    for rf_i, rf in enumerate(vocab["l_returnflag"]):
        for ls_i, ls in enumerate(vocab["l_linestatus"]):
            m = keep & (codes["l_returnflag"] == rf_i) & (codes["l_linestatus"] == ls_i)
            n = int(m.sum().item())
            if n == 0:
                continue
            qty, price, disc, tax = (c[k][m] for k in ("l_quantity", "l_extendedprice", "l_discount", "l_tax"))
            dp = price * (100 - disc)
            s_qty, s_price, s_dp = int(qty.sum().item()), int(price.sum().item()), int(dp.sum().item())
            s_ch, s_disc = int((dp * (100 + tax)).sum().item()), int(disc.sum().item())
            expect[(rf, ls)] = (s_qty, s_price, s_dp, s_ch, trunc_div(s_qty * 10**4, n), trunc_div(s_price * 10**4, n),
                                trunc_div(s_disc * 10**4, n), n)
    mt = tpch.to_device_table(gpu_ctx, lineitem)
    plan = tpch.q1_plan(tpch.Database(SF, None, None, mt))
    got = rows_of(plan.execute(gpu_ctx))
    assert "fused_scan_agg" in plan.last_strategy()
    assert len(got) == len(expect) == 4
    for r in got:
        assert tuple(raw_dec(x) for x in r[2:9]) + (r[9],) == expect[(r[0], r[1])], r
    mt._dev.free()


def _dec_lo_words(res, col):
    """Decimal128 result column -> its low 64-bit words (asserting that every value fits them)."""
    w = column_bytes_tensor(res, col)[0].view(torch.int64).view(-1, 2)
    assert bool(((w[:, 1] == 0) | ((w[:, 1] == -1) & (w[:, 0] < 0))).all().item())
    return w[:, 0]


def test_q3_sf10_every_group(gpu_ctx, lineitem):
    """EXACT per-group check at SF10: every (l_orderkey, revenue, o_orderdate, o_shippriority) row equals what torch's
    unique(return_inverse) + index_add_ give over the same generated columns (int64 arithmetic is exact at this size)."""
    cust = tpch.gen_customer(SF, device="cuda", columns=bench.QUERY_COLUMNS["q3"]["customer"])
    orders = tpch.gen_orders(SF, device="cuda", columns=bench.QUERY_COLUMNS["q3"]["orders"])
    d = tpch.days("1995-03-15")
    building = cust.cols["c_custkey"][cust.codes["c_mktsegment"] == cust.vocab["c_mktsegment"].index("BUILDING")]
    o = orders.cols
    o_keep = (o["o_orderdate"] < d) & torch.isin(o["o_custkey"], building)
    j1_keys, j1_date, j1_prio = o["o_orderkey"][o_keep], o["o_orderdate"][o_keep], o["o_shippriority"][o_keep]
    c = lineitem.cols
    l_keep = (c["l_shipdate"] > d) & torch.isin(c["l_orderkey"], j1_keys)
    lk = c["l_orderkey"][l_keep]
    val = c["l_extendedprice"][l_keep] * (100 - c["l_discount"][l_keep])
    exp_key, inv = torch.unique(lk, return_inverse=True)                       # sorted group keys
    exp_rev = torch.zeros(exp_key.numel(), dtype=torch.int64, device="cuda").index_add_(0, inv, val)
    srt, perm = torch.sort(j1_keys)
    pos = perm[torch.searchsorted(srt, exp_key)]
    exp_date, exp_prio = j1_date[pos], j1_prio[pos]
    tabs = {"customer": tpch.to_device_table(gpu_ctx, cust), "orders": tpch.to_device_table(gpu_ctx, orders),
            "lineitem": tpch.to_device_table(gpu_ctx, lineitem)}
    plan = tpch.q3_plan(tpch.Database(SF, tabs["customer"], tabs["orders"], tabs["lineitem"]))
    res = plan.execute_device(gpu_ctx)
    assert "fused_join_probe_agg" in plan.last_strategy()
    assert res.num_rows == exp_key.numel()
    names = [f.name for f in plan.schema]
    keys = column_bytes_tensor(res, names.index("l_orderkey"))[0].view(torch.int64)
    order = torch.argsort(keys)
    assert torch.equal(keys[order], exp_key)                                                        # every key exactly once
    assert torch.equal(_dec_lo_words(res, names.index("revenue"))[order], exp_rev)                  # every group's revenue
    assert torch.equal(column_bytes_tensor(res, names.index("o_orderdate"))[0].view(torch.int32)[order].to(torch.int64),
                       exp_date.to(torch.int64))
    assert torch.equal(column_bytes_tensor(res, names.index("o_shippriority"))[0].view(torch.int64)[order], exp_prio.to(torch.int64))
    res.free()
    for t in tabs.values():
        t._dev.free()


def test_groupby_1b_rows_100m_groups_every_group(gpu_ctx):
    """BASELINE.json configs[3] at full size, EXACT per group: SUM / COUNT / MIN / MAX(v) equal torch's index_add_ /
    bincount / scatter_reduce over unique(return_inverse) of the same keys; AVG(f) within 1e-12 relative (Float64: only
    the reduction order differs, north_star)."""
    rows, groups = 1_000_000_000, 100_000_000
    raw = bench.gen_groupby(rows, groups, "cuda")
    k, v, f = raw.cols["k"], raw.cols["v"], raw.cols["f"]
    exp_key, inv = torch.unique(k, return_inverse=True)
    g = exp_key.numel()
    exp_cnt = torch.bincount(inv, minlength=g)
    exp_sum = torch.zeros(g, dtype=torch.int64, device="cuda").index_add_(0, inv, v)
    exp_min = torch.full((g,), 2**62, dtype=torch.int64, device="cuda").scatter_reduce_(0, inv, v, "amin")
    exp_max = torch.full((g,), -2**62, dtype=torch.int64, device="cuda").scatter_reduce_(0, inv, v, "amax")
    exp_avg = torch.zeros(g, dtype=torch.float64, device="cuda").index_add_(0, inv, f) / exp_cnt.to(torch.float64)
    del inv
    mt = tpch.to_device_table(gpu_ctx, raw)
    del raw, k, v, f
    torch.cuda.empty_cache()
    plan = bench.groupby_plan(mt)
    res = plan.execute_device(gpu_ctx)
    assert "radix-partitioned" in plan.last_strategy(), plan.last_strategy()
    col = lambda i, dt: column_bytes_tensor(res, i)[0].view(dt)  # noqa: E731
    assert res.num_rows == g
    order = torch.argsort(col(0, torch.int64))
    assert torch.equal(col(0, torch.int64)[order], exp_key)                    # every key exactly once
    assert torch.equal(col(1, torch.int64)[order], exp_sum)                    # SUM(v) of every group
    assert torch.equal(col(2, torch.int64)[order], exp_cnt)                    # COUNT(v)
    assert torch.equal(col(3, torch.int64)[order], exp_min)                    # MIN(v)
    assert torch.equal(col(4, torch.int64)[order], exp_max)                    # MAX(v)
    got_avg = col(5, torch.float64)[order]
    assert bool(((got_avg - exp_avg).abs() <= 1e-12 * exp_avg.abs()).all().item())   # AVG(f): tolerance of north_star
    del order, got_avg
    res.free()
    mt._dev.free()
    gpu_ctx.release_cached_memory()      # ~80 GB of re-usable blocks: give them back for the tests that follow
    torch.cuda.empty_cache()
