"""Multi-rank parity of the NVLink halo exchange and the device all-reduce: tools/multi_gpu_check.py under torchrun with
2 ranks.  Each rank checks its brick against the single-rank oracle on the whole mesh.  With >= 2 GPUs one rank per
GPU (operator, BP5, BPS5); on a 1-GPU box BOTH ranks share the GPU (CUDA-IPC windows between two processes on one
device, kernels time-sliced): the same exchange code, correctness only, operator + a short BP5."""
import os
import subprocess
import sys

import pytest

from nekrs_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_operator_and_solves_match_single_rank_oracle():
    env = dict(os.environ, CHECK_N="7", CHECK_NEL="4,2,2")
    if lib.device_count() < 2:
        env.update(CHECK_SAME_GPU="1", CHECK_LIGHT="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0 and "MULTI_GPU_CHECK PASS" in p.stdout, p.stdout[-4000:]
