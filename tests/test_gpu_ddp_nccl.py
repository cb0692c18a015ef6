"""GPU x2: the data-parallel step over real NCCL: arena-averaged gradients == mean of the per-rank
gradients, and all 314 parameter tensors bit-identical on both ranks after captured steps.  Skipped with fewer than 2 GPUs."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [64])
def test_arena_all_reduce_matches_mean_of_rank_gradients_and_replicas_stay_in_sync(size):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29655", os.path.join(ROOT, "tests", "nccl_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=400, env=dict(os.environ, DUPL_TEST_SIZE=str(size)))
    assert out.returncode == 0, out.stderr[-3000:]
    r = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("RESULT ")][-1][7:])
    print(json.dumps(r))
    assert r["world"] == 2 and not r["missing"]
    assert r["grad_mean_worst_rel"] < 1e-5, r           # same kernels, same inputs: only the summation order of the average differs
    assert r["params_in_sync"] and r["params_checked"] == 314 and r["params_moved"] >= 300, r
    assert min(r["chunks"]) >= 2
