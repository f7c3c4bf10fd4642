"""The oracle (oracle/qref.py) pinned against every known-answer vector the reference's tests hold
for the hot path (tests/golden/reference_vectors.json).  CPU only."""
import pyarrow as pa
import pytest

from oracle import qref
from tests.cases import GOLDEN, check_rows, golden_cases, rows_of

CASES = list(golden_cases())


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_vector(case):
    name, plan, expected, ordered = case
    got = rows_of(qref.execute(plan))
    check_rows(name, got, expected, ordered)


def test_decimal_nested_result_type():
    # binary.rs:244-248: the result type is Decimal128(32, 4)
    name, plan, expected, _ = next(c for c in CASES if c[0] == "decimal_nested")
    out = qref.execute(plan)
    assert out[0].schema.field(0).type == pa.decimal128(32, 4)
    col = qref.from_arrow(out[0].column(0))
    assert [int(v) for v in col.vals] == GOLDEN["decimal_nested"]["expected_raw"]


def test_join_hash_map_update_vectors():
    for v in GOLDEN["join_hash_map"]["update"]:
        m = qref.JoinHashMap(len(v["hashes"]))
        m.update(enumerate(v["hashes"]), v["delete_offset"])
        assert {str(k): val for k, val in m.map.items()} == v["map"], v["source"]
        assert m.next == v["next"], v["source"]
        assert m.is_distinct() == v["distinct"], v["source"]


def test_join_hash_map_delete_offset():
    # hash_join.rs:763-779
    m = qref.JoinHashMap(3)
    m.update(enumerate([100, 200, 300]), 2)
    assert m.map == {100: 1, 200: 2, 300: 3} and m.next == [0, 0, 0]


def test_join_hash_map_match_vectors():
    for v in GOLDEN["join_hash_map"]["matches"]:
        m = qref.JoinHashMap(len(v["build"]))
        m.update(enumerate(v["build"]), 0)
        inp, mat = m.get_matches_indices(v["probe"])
        assert inp == v["input_indices"] and mat == v["match_indices"], v["source"]


def test_reverse_build_gives_ascending_chains():
    # hash_join.rs:166: build inserts rows in REVERSE so every chain lists build rows ascending
    hashes = [100, 200, 100, 300, 100]
    m = qref.JoinHashMap(len(hashes))
    m.update(reversed(list(enumerate(hashes))), 0)
    inp, mat = m.get_matches_indices([100])
    assert mat == [0, 2, 4]


def test_cast_literal_vectors():
    from qurious_b200.datatypes import ScalarValue
    from qurious_b200.physical.expr import CastExpr, Literal
    one = pa.record_batch([pa.array([0])], names=["x"])
    g = GOLDEN["cast_literals"]
    for f, raw in g["float_to_decimal_15_2"]:
        c = qref.evaluate(CastExpr(Literal(ScalarValue.Float64(f)), pa.decimal128(15, 2)), one)
        assert int(c.vals[0]) == raw
    for i, raw in g["int_to_decimal_15_2"]:
        assert int(qref.evaluate(CastExpr(Literal(ScalarValue.Int64(i)), pa.decimal128(15, 2)), one).vals[0]) == raw
    for i, raw in g["int_to_decimal_20_0"]:
        assert int(qref.evaluate(CastExpr(Literal(ScalarValue.Int64(i)), pa.decimal128(20, 0)), one).vals[0]) == raw
    for s, days in g["utf8_to_date32"]:
        assert int(qref.evaluate(CastExpr(Literal(ScalarValue.Utf8(s)), pa.date32()), one).vals[0]) == days


def test_cast_cross_check_with_arrow_cpp():
    """Independent cross-check of the cast restatement against Arrow C++ (pyarrow), SURVEY 8c."""
    import pyarrow.compute as pc
    # (exact .5 ties are excluded: arrow-rs uses f64::round = half away from zero, Arrow C++ does not)
    f = pa.array([0.049999999999999996, 0.06999999999999999, 0.124, -0.126, 1.005, 2.5, 123456.789])
    got = qref.cast_col(qref.from_arrow(f), pa.decimal128(15, 2))
    exp = pc.cast(f, pa.decimal128(15, 2))
    assert [int(v) for v in got.vals] == [int(x.as_py().scaleb(2)) for x in exp]
    s = pa.array(["1998-09-02", "1992-01-01", "2000-02-29"])
    got = qref.cast_col(qref.from_arrow(s), pa.date32())
    assert [int(v) for v in got.vals] == pc.cast(s, pa.date32()).cast(pa.int32()).to_pylist()
