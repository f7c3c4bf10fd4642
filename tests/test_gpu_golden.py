"""GPU operators (through the C ABI) against the reference's known-answer vectors and the oracle."""
import pyarrow as pa
import pytest

from oracle import qref
from tests.cases import check_rows, golden_cases, rows_of

pytestmark = pytest.mark.gpu
CASES = list(golden_cases())


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_gpu_matches_reference_vector(case, gpu_ctx):
    name, plan, expected, ordered = case
    got_batches = plan.execute(gpu_ctx)
    got = rows_of(got_batches)
    check_rows(name, got, expected, ordered)
    # and the oracle on the same plan object
    ref_batches = qref.execute(plan)
    check_rows(name + " (vs oracle)", got, rows_of(ref_batches), ordered)
    if ref_batches and got_batches:
        assert got_batches[0].schema.types == ref_batches[0].schema.types, name


def test_decimal_nested_type_and_raw(gpu_ctx):
    name, plan, expected, _ = next(c for c in CASES if c[0] == "decimal_nested")
    out = plan.execute(gpu_ctx)
    assert out[0].schema.field(0).type == pa.decimal128(32, 4)
    assert [int(v) for v in qref.from_arrow(out[0].column(0)).vals] == [259873950]


def test_join_schema_matches_host_mirror(gpu_ctx):
    import ctypes
    from qurious_b200 import _lib
    from pyarrow.cffi import ffi
    for name, plan, _, _ in CASES:
        if not name.startswith(("hash_join:", "join:")):
            continue
        ctx, h, keep = plan._native(gpu_ctx)
        try:
            cs = ffi.new("struct ArrowSchema*")
            ctx.check(ctx.lib.qgpu_plan_schema(h, _lib.addr(cs)))
            got = pa.Schema._import_from_c(_lib.addr(cs))
        finally:
            plan._free(ctx, keep)
        assert got.equals(plan.schema, check_metadata=True), name
