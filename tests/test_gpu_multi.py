"""Multi-rank parity of the CUDA library under the product's own multi-rank schedule (Engine.step with a process
group): particles sharded, grids all-reduced, spectral solve on kx slabs, slabs all-gathered.

With two or more GPUs on the box the ranks get a GPU each and talk NCCL.  With ONE GPU (the driver's test lease) the two
ranks share it and talk gloo -- NCCL refuses two ranks on one device -- so the same schedule and the same kernels are
exercised either way and nothing here is skipped."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import chimera_b200.fimera as f

    return f.device_count()


CASES = [("real_m2", 1, 0), ("real_m2", 0, 0), ("env_m3", 1, 0), ("real_m2", 1, 1),
         ("static_m2", 1, 0),   # 'StaticKick' on kx slabs: moments all-reduced, quasi-static field per slab
         ("static_m2", 0, 1),   # ... with the space-charge demo's 'Staged' frame, spectral solve replicated
         ("real_m2", 1, 2),     # the LPA window (frame_act) across ranks: slab damp_field, sharded injection, cull
         ("real_m3", 0, 2)]


@pytest.mark.parametrize("name,slab,window", CASES)
def test_two_ranks(name, slab, window):
    backend = "nccl" if _ngpu() >= 2 else "gloo"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(HERE, "dist_gpu_worker.py"), name, str(slab), str(window),
           backend]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-4000:]
