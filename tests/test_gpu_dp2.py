"""Two-GPU data-parallel equivalence on real devices (NCCL): sharded `Trainer.step` with the bucketed, overlapped all-reduce
against a single-process step on the whole batch.  Skipped where fewer than two GPUs are visible (the CPU / gloo version of
the same property is tests/test_dp_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

from tests.helpers import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharded_step_equals_single_process():
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp2_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "dp2 worst per-block relative L2 error" in r.stdout
