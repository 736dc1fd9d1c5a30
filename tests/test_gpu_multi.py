"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches tools/multi_gpu_check.py under
torchrun with 2 ranks.  Each rank checks its brick against the single-rank oracle on the whole mesh."""
import os
import subprocess
import sys

import pytest

from nekrs_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_gpu_operator_and_solves_match_single_rank_oracle():
    if lib.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, CHECK_N="7", CHECK_NEL="4,2,2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0 and "MULTI_GPU_CHECK PASS" in p.stdout, p.stdout[-4000:]
