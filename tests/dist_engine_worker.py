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


def main(window, overlap, name="real_m2"):
    from cpu_engine import CpuEngine, reference_run
    from util import SETUPS, assert_close, plasma, seed_fields
    from chimera_b200 import sharding
    from chimera_b200.solver_setup import SolverSetup

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    S = SolverSetup(copy.deepcopy(SETUPS[name]))
    x, p, w = plasma(S, 2, 2, 91)
    xi, pi_, wi = plasma(S, 2, 2, 97)
    eg0 = seed_fields(S, 92)
    nsteps = 6  # crosses a re-binning step (Xchunked = (4, 3)): fused and unfused branches of the schedule
    win = (0.5 * 0.37 * S.Args["dt"],) * 2 if window else (0.0, 0.0)
    ions = "SpaceCharge" in S.Args.get("Features", ())
    species = [(x, p, w, {})] + ([(xi, 0 * pi_, -wi, dict(charge=1.0, mass=1886.0, still=True))] if ions else [])
    ref = reference_run(S, species, eg0, nsteps, win)

    eng = CpuEngine(S, group=True)
    assert eng.slab and eng.world == 2
    eng.overlap = bool(overlap)
    lo, hi = sharding.particle_range(x.shape[1], rank, world)
    ilo, ihi = sharding.particle_range(xi.shape[1], rank, world)
    eng.add_species(x[:, lo:hi], p[:, lo:hi], w[lo:hi])
    if ions:
        eng.add_species(xi[:, ilo:ihi], 0 * pi_[:, ilo:ihi], -wi[ilo:ihi], charge=1.0, mass=1886.0, still=True)
    eng.upload("EG_fb", eg0)
    if window:
        eng.set_window(0.37, staged=True)
    eng.make_halfstep(px0=(0.0,) * len(species), background=ions)
    eng.step(2)
    eng.step(nsteps - 2)
    from util import carrier_tol

    tol = carrier_tol(S, 1e-11)
    assert_close(eng.download("EG_fb"), ref.EG_fb[eng.rows], tol, "EG_fb slab")
    assert_close(eng.download("EB"), ref.EB, tol, "EB")
    assert_close(eng.download("J"), ref.J, 20 * tol if S.env else tol, "J")
    if ions:
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
        print("OK window", window, "overlap", overlap, name)


def main_lpa():
    """the LPA moving-window run recorded from the reference's own driver (tests/golden/real_m2_lpa.npz): species that
    start empty, Engine.frame_act every 3 steps with the added particles sharded over the ranks"""
    import json

    from scipy.spatial import cKDTree

    from cpu_engine import CpuEngine
    from util import SETUPS, assert_close
    from chimera_b200 import sharding
    from chimera_b200.solver_setup import SolverSetup

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    z = np.load(os.path.join(ROOT, "tests", "golden", "real_m2_lpa.npz"))
    meta = json.loads(str(z["cfg"]))
    case, nsteps = meta["case"], meta["nsteps"]
    adds = {}
    for k, (st, si) in enumerate(zip(z["add_steps"], z["add_species"])):
        x, p, w = z["add%d_coords" % k], z["add%d_momenta" % k], z["add%d_weights" % k]
        lo, hi = sharding.particle_range(x.shape[1], rank, world)
        adds.setdefault(int(st), {})[int(si)] = (x[:, lo:hi], p[:, lo:hi], w[lo:hi])
    wind = {"shiftX": float(z["shiftX"]), "AbsorbLayer": case["wind"]["AbsorbLayer"], "Steps": case["wind"]["Steps"]}
    S = SolverSetup(copy.deepcopy(SETUPS[case["setup"]]))
    e0 = np.zeros((3, 0), order="F")
    eng = CpuEngine(S, group=True, sort_every=0)
    assert eng.slab
    eng.add_species(e0, e0, np.zeros(0))
    eng.add_species(e0, e0, np.zeros(0), charge=1.0, mass=1886.0, still=True)
    eng.upload("EG_fb", z["in_EG_fb"])
    eng.make_halfstep(px0=(0.0, 0.0), background=True)
    assert_close(eng.download("EG_fb"), z["h_EG_fb"][eng.rows], 1e-11, "EG_fb after make_halfstep")
    import torch

    for i in range(1, nsteps + 1):
        if i % wind["Steps"] == 0:
            eng.frame_act(wind, add=adds.get(i), background=True)
        eng.step(1)
        pre = "s%d" % i
        if pre + "_EG_fb" in z.files:
            assert_close(eng.download("EG_fb"), z[pre + "_EG_fb"][eng.rows], 2e-11, "EG_fb slab at step %d" % i)
            for k in ("J", "Rho", "BckGrndRho", "EB"):
                assert_close(eng.download(k), z[pre + "_" + k], 2e-11, "%s at step %d" % (k, i))
            xs, xh, ps, ws = eng.particles(0)
            dist_, perm = cKDTree(z[pre + "_coords"].T).query(xs.T)  # this rank's particles are a subset of the recorded set
            assert np.unique(perm).size == perm.size and dist_.max() <= 1e-9
            assert_close(ps, z[pre + "_momenta"][:, perm], 2e-11, "momenta at step %d" % i)
            cnt = torch.tensor([float(ws.size), float(eng.count(1))], dtype=torch.float64)
            dist.all_reduce(cnt)
            assert int(cnt[0]) == z[pre + "_weights"].size and int(cnt[1]) == z[pre + "_ion_weights"].size
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("OK lpa window")


if __name__ == "__main__":
    if sys.argv[1] == "lpa":
        main_lpa()
    else:
        main(int(sys.argv[1]), int(sys.argv[2]), *sys.argv[3:4])
