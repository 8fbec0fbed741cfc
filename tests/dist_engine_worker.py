"""Worker of tests/test_dist_gloo.py::test_engine_schedule_*: one rank of a world_size-2 gloo job (CPU) that runs the
PRODUCT's multi-rank host logic -- Engine.make_halfstep / Engine.step of chimera_b200/engine.py, unchanged -- on the
CPU stand-in of the library (tests/cpu_engine.py) and compares every rank's state with the single-process reference
sequence.  argv: window (0 | 1), overlap (0 | 1)."""
import copy
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(window, overlap):
    from cpu_engine import CpuEngine, reference_run
    from util import SETUPS, assert_close, plasma, seed_fields
    from chimera_b200 import sharding
    from chimera_b200.solver_setup import SolverSetup

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    S = SolverSetup(copy.deepcopy(SETUPS["real_m2"]))
    x, p, w = plasma(S, 2, 2, 91)
    xi, pi_, wi = plasma(S, 2, 2, 97)
    eg0 = seed_fields(S, 92)
    nsteps = 6  # crosses a re-binning step (Xchunked = (4, 3)): fused and unfused branches of the schedule
    win = (0.5 * 0.37 * S.Args["dt"],) * 2 if window else (0.0, 0.0)
    ref = reference_run(S, [(x, p, w, {}), (xi, 0 * pi_, -wi, dict(charge=1.0, mass=1886.0, still=True))], eg0, nsteps, win)

    eng = CpuEngine(S, group=True)
    assert eng.slab and eng.world == 2
    eng.overlap = bool(overlap)
    lo, hi = sharding.particle_range(x.shape[1], rank, world)
    ilo, ihi = sharding.particle_range(xi.shape[1], rank, world)
    eng.add_species(x[:, lo:hi], p[:, lo:hi], w[lo:hi])
    eng.add_species(xi[:, ilo:ihi], 0 * pi_[:, ilo:ihi], -wi[ilo:ihi], charge=1.0, mass=1886.0, still=True)
    eng.upload("EG_fb", eg0)
    if window:
        eng.set_window(0.37, staged=True)
    eng.make_halfstep(px0=(0.0, 0.0), background=True)
    eng.step(2)
    eng.step(nsteps - 2)
    tol = 1e-11
    assert_close(eng.download("EG_fb"), ref.EG_fb[eng.rows], tol, "EG_fb slab")
    assert_close(eng.download("EB"), ref.EB, tol, "EB")
    assert_close(eng.download("J"), ref.J, tol, "J")
    assert_close(eng.download("Rho"), ref.Rho, tol, "Rho")
    xs, xh, ps, ws = eng.particles(0)
    order = np.argsort(ref.sp[0].weights)
    idx = order[np.searchsorted(ref.sp[0].weights[order], ws)]
    assert np.array_equal(ref.sp[0].weights[idx], ws)
    assert_close(ps, ref.sp[0].momenta[:, idx], tol, "momenta")
    assert_close(xs, ref.sp[0].coords[:, idx], tol, "coords")
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("OK window", window, "overlap", overlap)


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]))
