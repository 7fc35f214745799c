"""Two ranks, two GPUs, the library's own NCCL all-gather: every rank's slice of the gathered top-K buffer against the oracle
(tests/mgpu_worker.py does the work under torchrun).  Skipped on a single-GPU box; tests/test_sharding_gloo.py covers the host logic on CPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gathered_records_of_every_rank_match_the_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ)
    env.setdefault("MASTER_ADDR", "127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "0 mismatches" in r.stdout
