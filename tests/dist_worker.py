"""Worker of tests/test_dist_gloo.py: one rank of a world_size-2 gloo job (CPU).

Exercises the host-side multi-GPU logic of SURVEY.md section 8e without a GPU:
  (1) particles sharded over the ranks, deposited grids all-reduced, background charge added once;
  (2) the spectral update restricted to this rank's kx slab of mirror pairs (chimera_b200/sharding.py),
      slabs all-gathered afterwards.
Kernels underneath: the CPU oracle (1) and the numpy restatement (2).  Writes rank-local results to
the directory given as argv[1]; the test compares them with a single-process run."""
import copy
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(out_dir):
    from oracle import fimera as ofim
    from oracle import np_ref
    from pic_ref import RefRun, RefSpecies
    from util import SETUPS, plasma, seed_fields
    from chimera_b200 import sharding
    from chimera_b200.solver_setup import SolverSetup

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    S = SolverSetup(copy.deepcopy(SETUPS["real_m2"]))
    x, p, w = plasma(S, 2, 2, 41)
    xi, pi_, wi = plasma(S, 2, 2, 47)
    lo, hi = sharding.particle_range(x.shape[1], rank, world)
    ilo, ihi = sharding.particle_range(xi.shape[1], rank, world)
    sp = [RefSpecies(x[:, lo:hi], p[:, lo:hi], w[lo:hi]),
          RefSpecies(xi[:, ilo:ihi], 0 * pi_[:, ilo:ihi], -wi[ilo:ihi], charge=1.0, mass=1886.0, still=True)]

    def allreduce_f(a):  # Fortran-ordered complex grid, in place
        flat = np.ascontiguousarray(a.ravel(order="K").view(np.float64))
        t = torch.from_numpy(flat)
        dist.all_reduce(t)
        a.ravel(order="K")[:] = t.numpy().view(np.complex128)

    run = RefRun(ofim, S, sp, background=True, reduce=allreduce_f, rank=rank)
    run.EG_fb[:] = seed_fields(S, 42)
    run.make_halfstep(px0=(0.0, 0.0))
    run.make_step()
    run.make_step()
    np.savez(os.path.join(out_dir, "particles_%d.npz" % rank), EG_fb=run.EG_fb, J=run.J, Rho=run.Rho,
             momenta=run.sp[0].momenta, weights=run.sp[0].weights)

    # ---- (2) kx-slab sharded spectral update on a fixed input
    a = S.Args
    rng = np.random.default_rng(5)  # same on every rank
    shp = S.shape_fb
    J = rng.standard_normal(shp + (3,)) + 1j * rng.standard_normal(shp + (3,))
    g0 = rng.standard_normal(shp + (3,)) + 1j * rng.standard_normal(shp + (3,))
    g1 = rng.standard_normal(shp + (3,)) + 1j * rng.standard_normal(shp + (3,))
    EG = rng.standard_normal(shp + (6,)) + 1j * rng.standard_normal(shp + (6,))
    rows = sharding.kx_slab_rows(shp[0], rank, world)
    np_ref.MIRROR_SHIFT = sharding.mirror_shift(rank, world)
    Dp, Dm, kx = a["FBDiff"]
    Js, g0s, g1s, EGs = J[rows], g0[rows], g1[rows], EG[rows]
    for _ in range(3):
        gd = np_ref.fb_graddiv(Js, Dp, Dm, kx[rows])
        Js = np_ref.poiss_corr(Js, gd, g0s, g1s, a["dt_inv"], a["PoissFact"][rows])
    EGs = np_ref.maxwell_push_with_spchrg(EGs, Js, g0s, g1s, S.PSATD_E[rows], S.PSATD_G[rows])
    Bs = np_ref.fb_rot(EGs[..., 3:], Dp, Dm, kx[rows]) * a["PoissFact"][rows][..., None]
    np_ref.MIRROR_SHIFT = 0
    out = np.ascontiguousarray(np.concatenate((EGs, Bs), axis=-1))
    gathered = [torch.zeros(out.shape, dtype=torch.complex128) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(out))
    full = np.zeros(shp + (9,), dtype=complex)
    for r in range(world):
        full[sharding.kx_slab_rows(shp[0], r, world)] = gathered[r].numpy()
    if rank == 0:
        np.savez(os.path.join(out_dir, "slab.npz"), full=full)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
