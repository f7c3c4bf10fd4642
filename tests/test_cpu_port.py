"""The C++ CPU port (oracle/qref_cpu.cpp, the reported cpu_baseline) against the numpy oracle."""
from oracle import cpu_port, qref
from qurious_b200 import tpch
from tests.cases import rows_of


def _raw(v):
    return int(v.scaleb(-v.as_tuple().exponent)) if v is not None else None


def test_q6_port_matches_oracle():
    db = tpch.generate(0.005)
    ref = rows_of(qref.execute(tpch.q6_plan(db)))
    got = cpu_port.q6(db.lineitem.data)
    assert got["rows"] > 0
    assert _raw(ref[0][0]) == got["revenue_raw"]


def test_q1_port_matches_oracle():
    db = tpch.generate(0.005)
    ref = rows_of(qref.execute(tpch.q1_plan(db)))
    got = cpu_port.q1(db.lineitem.data)
    want = sorted((r[0], r[1], *[_raw(x) for x in r[2:9]], r[9]) for r in ref)
    assert sorted(got) == want


def test_q3_port_matches_oracle():
    import datetime
    db = tpch.generate(0.005)
    ref = rows_of(qref.execute(tpch.q3_plan(db)))
    got = cpu_port.q3(db.customer.data, db.orders.data, db.lineitem.data)
    assert len(got) > 0
    epoch = datetime.date(1970, 1, 1)
    want = sorted((r[0], _raw(r[1]), (r[2] - epoch).days, r[3]) for r in ref)
    assert sorted(got) == want


def test_groupby_port_matches_oracle():
    import numpy as np
    import pyarrow as pa
    import sys
    sys.path.insert(0, ".")
    import bench
    from qurious_b200.physical.plan import MemoryTable
    rng = np.random.default_rng(4)
    n = 20000
    schema = pa.schema([("k", pa.int64()), ("v", pa.int64()), ("f", pa.float64())])
    b = pa.record_batch([pa.array(rng.integers(0, 3000, n) * 7919 - 10**9), pa.array(rng.integers(-10**6, 10**6, n)),
                         pa.array(rng.random(n))], schema=schema)
    t = MemoryTable.try_new(schema, [b])
    ref = sorted(rows_of(qref.execute(bench.groupby_plan(t))))
    got = sorted(cpu_port.groupby([b]))
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert g[:5] == tuple(r[:5])
        assert abs(g[5] - r[5]) <= 1e-12 * max(1.0, abs(r[5]))
