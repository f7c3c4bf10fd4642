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
