"""Late materialisation, measured: the reference copies EVERY column of the surviving rows at every operator
(filter_record_batch, memory.rs:90-92; `take` of all build + probe columns per join, utils/batch.rs:18-61) -- the oracle
restates exactly that, so the bytes of its intermediate batches are the bytes the reference materialises.  The GPU operators
exchange index vectors and gather payload columns once, when a consumer needs them: qgpu_counter reports the bytes of all
device temporaries of an execution and the payload bytes its gathers wrote."""
import pytest

from oracle import qref
from qurious_b200 import tpch
from tests.cases import check_rows, rows_of

pytestmark = pytest.mark.gpu


def reference_materialised_bytes(plan) -> int:
    """Bytes of every RecordBatch the reference's operators produce below the root (filters, joins, aggregates)."""
    total = 0

    def walk(node):
        nonlocal total
        for c in (node.children() or []):
            total += sum(b.nbytes for b in qref.execute(c))
            walk(c)
    walk(plan)
    return total


def test_q3_all_17_36_columns_versus_index_vectors(gpu_ctx):
    db = tpch.generate(0.01)            # the full 9 / 10 / 17-column tables, like the reference's
    plan = tpch.q3_plan(db)
    ref_bytes = reference_materialised_bytes(plan)
    plan.execute(gpu_ctx)                                   # upload + analysis
    a0, g0 = gpu_ctx.counter("alloc_bytes"), gpu_ctx.counter("gather_bytes")
    got = plan.execute(gpu_ctx)
    alloc, gathered = gpu_ctx.counter("alloc_bytes") - a0, gpu_ctx.counter("gather_bytes") - g0
    check_rows("q3", rows_of(got), rows_of(qref.execute(plan)), ordered=False)
    print(f"\nQ3 SF0.01: reference materialises {ref_bytes / 1e6:.1f} MB of intermediate batches; GPU temporaries {alloc / 1e6:.2f} MB, "
          f"of which gathered payload {gathered / 1e6:.2f} MB")
    assert gathered * 20 < ref_bytes and alloc * 3 < ref_bytes
