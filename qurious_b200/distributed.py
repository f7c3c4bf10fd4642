"""Multi-GPU execution of aggregates over row-range shards: one process per GPU, NCCL through
torch.distributed for the plumbing.  The reference is single-process, so nothing here mirrors a reference file;
the merge itself is exact and happens on the GPU inside libqgpu (csrc/shard.cu):

    every rank:  partial state of (Projection* <- Aggregate <- Scan(shard))      qgpu_plan_partial_state
    NCCL:        all-gather of the fixed-size state blocks                       dist.all_gather_into_tensor
    every rank:  merge the gathered states + finalise (AVG division, order...)   qgpu_plan_execute_merged

No compute happens in Python.  `row_offset` is the global index of the shard's first row: it keeps the
first-occurrence output order identical to a single-GPU run over the whole table.
"""
from __future__ import annotations

import ctypes
from typing import Callable, List, Optional

import pyarrow as pa
import torch

from . import _lib


def shard_range(total_rows: int, rank: int, world: int):
    """Contiguous row range [lo, hi) of `rank` (rank r gets rows [r*N/G, (r+1)*N/G))."""
    return (total_rows * rank) // world, (total_rows * (rank + 1)) // world


class ShardedAggregate:
    """plan: (Projection|Filter)* <- HashAggregate/NoGroupingAggregate <- ... over THIS rank's shard."""

    def __init__(self, ctx: _lib.Context, plan, row_offset: int, world: int, max_groups: int = 64,
                 all_gather: Optional[Callable[[torch.Tensor, torch.Tensor], None]] = None):
        self.ctx, self.plan, self.row_offset, self.world, self.max_groups = ctx, plan, int(row_offset), int(world), int(max_groups)
        _, self.h, _ = plan._native_cached(ctx)
        n = ctypes.c_int64()
        ctx.check(ctx.lib.qgpu_plan_state_bytes(self.h, self.max_groups, ctypes.byref(n)))
        self.state_bytes = n.value
        dev = torch.device("cuda", ctx.device)
        self.state = torch.zeros(self.state_bytes, dtype=torch.uint8, device=dev)
        self.gathered = torch.zeros(self.state_bytes * self.world, dtype=torch.uint8, device=dev)
        self.stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)
        if all_gather is None:
            import torch.distributed as dist

            def all_gather(out, inp):
                dist.all_gather_into_tensor(out, inp)
        self.all_gather = all_gather
        torch.cuda.synchronize(dev)  # the zero-fills above ran on torch's stream

    def partial(self) -> torch.Tensor:
        """Run the shard-local part; returns the device state block (valid until the next call)."""
        self.ctx.check(self.ctx.lib.qgpu_plan_partial_state(self.h, self.row_offset, self.max_groups,
                                                            self.state.data_ptr(), self.state_bytes))
        return self.state

    def merge(self, gathered: torch.Tensor, n_states: int) -> List[pa.RecordBatch]:
        out = _lib.new_stream()
        self.ctx.check(self.ctx.lib.qgpu_plan_execute_merged(self.h, gathered.data_ptr(), n_states, self.max_groups,
                                                             _lib.addr(out)))
        self.plan._record_stats(self.ctx, self.h)
        return _lib.read_stream(self.ctx, out)

    def execute(self) -> List[pa.RecordBatch]:
        self.partial()
        with torch.cuda.stream(self.stream):       # the collective is ordered after the library's stream work
            self.all_gather(self.gathered, self.state)
        return self.merge(self.gathered, self.world)
