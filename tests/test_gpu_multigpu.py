"""Real-NCCL correctness of every sharded protocol: when the box has at least two GPUs, a 2-rank torchrun job
(scripts/check_multigpu.py) must find sharded Q1 / Q6 (fused peer exchange, asynchronous, and the three-call protocol over
the library's own NCCL all-gather), Q3 (broadcast join) and the exchange group-by equal to the single-table plans."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_nccl_job_equals_single_gpu_plans():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs (run with gpurun --gpus 2)")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "check_multigpu.py")]
    env = dict(os.environ, CHECK_SF="0.2")
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    tail = r.stdout[-4000:]
    assert r.returncode == 0 and "MULTIGPU CHECK OK" in r.stdout, tail
