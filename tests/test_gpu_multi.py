"""Multi-GPU (NCCL) parity: needs at least 2 CUDA devices on the box (skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import chimera_b200.fimera as f

    return f.device_count()


@pytest.mark.parametrize("name,slab,window", [("real_m2", 1, 0), ("real_m2", 0, 0), ("env_m3", 1, 0), ("real_m2", 1, 1)])
def test_two_ranks_nccl(name, slab, window):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(HERE, "dist_gpu_worker.py"), name, str(slab), str(window)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-4000:]
