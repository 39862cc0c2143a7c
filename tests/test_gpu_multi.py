"""-m gpu multi-rank tests (skipped with fewer than 2 GPUs): one process per GPU over NCCL, spawned with
torch.distributed.run; the checks themselves live in tests/mp_gpu_worker.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4])
def test_multi_rank_build_and_partitioned_message_passing(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 36000 + (os.getpid() % 2000) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    errs = [ln for ln in r.stderr.splitlines() if "Error" in ln or "differ" in ln]
    assert r.returncode == 0 and ("MP_GPU_OK world=%d" % world) in r.stdout, "\n".join(errs[-12:]) or r.stderr[-3000:]
