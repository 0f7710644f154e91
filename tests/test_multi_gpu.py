"""N > 1 parity on real GPUs: launches tests/multi_gpu_check.py under torchrun with two ranks (NCCL).  Skips itself on a
box with fewer than two GPUs; the host-side sharding / gather logic is covered on CPU by tests/test_shard_logic.py (gloo)."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count() -> int:
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.gpu
def test_two_rank_gather_parity():
    if _gpu_count() < 2:
        pytest.skip("needs at least two GPUs")
    env = dict(os.environ, PYTHONPATH=REPO, TZ="UTC")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(REPO, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=REPO)
    assert r.returncode == 0 and "[multi_gpu_check] ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
