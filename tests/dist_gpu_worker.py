"""Worker of tests/test_gpu_multi.py: one rank of an N-rank job on the CUDA library.  Particles sharded over the ranks,
deposited grids all-reduced, spectral solve sharded by kx slab, backward-transformed slabs all-gathered.
Compares every rank's state with the single-process oracle sequence and exits non-zero on a mismatch.

Backend "nccl": one GPU per rank.  Backend "gloo": the ranks SHARE GPU 0 (NCCL refuses two ranks on one device) -- the
same Engine schedule and the same kernels, only the bytes of the collectives travel through gloo; this is what runs when
the test box has a single GPU.
window: 0 none, 1 a 'Staged' frame every step, 2 the LPA window (frame_act: damp, move, add plasma, cull) between steps."""
import copy
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(name, slab, window=0, backend="nccl"):
    from oracle import fimera as ofim
    from pic_ref import RefRun, RefSpecies
    from util import SETUPS, TOL, assert_close, carrier_tol, plasma, seed_fields
    from chimera_b200 import sharding
    from chimera_b200.engine import Engine
    from chimera_b200.solver_setup import SolverSetup

    local = int(os.environ.get("LOCAL_RANK", 0)) if backend == "nccl" else 0
    torch.cuda.set_device(local)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from chimera_b200 import _lib

    _lib.load().chimera_set_device(local)
    S = SolverSetup(copy.deepcopy(SETUPS[name]))
    x, p, w = plasma(S, 2, 2, 71)
    xi, pi_, wi = plasma(S, 2, 2, 77)
    ions = "SpaceCharge" in S.Args.get("Features", ())
    static = "StaticKick" in S.Args.get("Features", ())
    if static:
        p[0] += 50.0  # the space-charge demo's beam (chimera_main.py:118-125 rebuilds the field from its mean momentum)
    eg0 = seed_fields(S, 72) * (0.0 if static else 1.0)
    sp = [RefSpecies(x, p, w)]
    if ions:
        sp.append(RefSpecies(xi, 0 * pi_, -wi, charge=1.0, mass=1886.0, still=True))
    # the reference sequence moves the window by editing its solver dictionary: give it its own
    ref = RefRun(ofim, SolverSetup(copy.deepcopy(SETUPS[name])) if window else S, sp, background=ions)
    ref.EG_fb[:] = eg0
    px0 = ((50.0 if static else 0.0),) * len(sp)
    if window == 1:  # a 'Staged' frame moving every step (chimera_main.py:40-51, 83-87)
        ref.window = (0.5 * 0.37 * S.Args["dt"],) * 2
    ref.make_halfstep(px0=px0)
    wind = {"shiftX": 4 * S.Args["dx"], "AbsorbLayer": 24, "Features": ()}

    def fresh(k):  # plasma entering on the right after the k-th window move, the same on every rank
        rng = np.random.default_rng(700 + k)
        a = ref.a
        xr = a["Xgrid"][-1] + wind["shiftX"]
        xa = np.asfortranarray(np.vstack((xr - rng.random(300) * wind["shiftX"], (rng.random((2, 300)) - 0.5) * 1.2 * a["Rgrid"].max())))
        return xa, np.asfortranarray(rng.standard_normal((3, 300)) * 0.05), -np.abs(rng.random(300)) * 1e-3 * (1 + 1e-6 * np.arange(300))

    adds = []
    if window == 2:  # the LPA window (chimera_main.py:250-304) twice, steps in between
        for k in range(2):
            ref.make_step()
            xa, pa, wa = fresh(k)
            adds.append((xa, pa, wa))
            add = {0: (xa, pa, wa)}
            if ions:
                add[1] = (xa.copy(order="F"), np.zeros_like(pa), -wa)
            ref.frame_act(wind, add)
        ref.make_step()
    else:
        for _ in range(3):
            ref.make_step()

    eng = Engine(S, group=True, slab=bool(slab))
    lo, hi = sharding.particle_range(x.shape[1], rank, world)
    eng.add_species(x[:, lo:hi], p[:, lo:hi], w[lo:hi])
    if ions:
        ilo, ihi = sharding.particle_range(xi.shape[1], rank, world)
        eng.add_species(xi[:, ilo:ihi], 0 * pi_[:, ilo:ihi], -wi[ilo:ihi], charge=1.0, mass=1886.0, still=True)
    eng.upload("EG_fb", eg0)
    if window == 1:
        eng.set_window(0.37, staged=True)
    eng.make_halfstep(px0=px0, background=ions)
    if window == 2:
        for k in range(2):
            eng.step(1)
            xa, pa, wa = adds[k]
            alo, ahi = sharding.particle_range(xa.shape[1], rank, world)  # every rank injects its share
            add = {0: (xa[:, alo:ahi], pa[:, alo:ahi], wa[alo:ahi])}
            if ions:
                add[1] = (xa[:, alo:ahi].copy(order="F"), np.zeros_like(pa[:, alo:ahi]), -wa[alo:ahi])
            eng.frame_act(wind, add, background=ions)
    else:
        eng.step(2)
    xs, xh, ps, ws = eng.particles(0)
    eg = eng.download("EG_fb")
    if window or static:  # windows and the static-kick schedule live in the step schedule, not in the host-buffer entry point
        eng.step(1)
        xs, xh, ps, ws = eng.particles(0)
        eg = eng.download("EG_fb")
        n = ws.size
    else:
        # third step through the host-buffer entry point (begin / all-reduce / mid / all-gather / end)
        g = eng.download("gradRho_fb_nxt") if eng.cfg.space_charge else None
        n = eng.step_host(xs, xh, ps, ws, eg, g)
    tol = carrier_tol(S, 4 * TOL)
    rows = eng.rows if eng.slab else slice(None)
    assert_close(eg, ref.EG_fb[rows], tol, "EG_fb")
    assert_close(eng.download("EG_fb"), ref.EG_fb[rows], tol, "EG_fb (device)")
    assert_close(eng.download("EB"), ref.EB, tol, "EB")
    if eng.colflow and (window or static):
        # column-block dataflow: the ranks' J is reduce-scattered, the sum only exists as column blocks -- sum a copy
        # (without a window the last step above went through step_host, which all-reduces J in place)
        jt = eng.device_tensor("J").clone()
        dist.all_reduce(jt)
        jsum = jt.cpu().numpy().view(np.complex128).reshape(ref.J.shape, order="F")
    else:
        jsum = eng.download("J")
    assert_close(jsum, ref.J, 20 * tol if S.env else tol, "J")
    # this rank's particles are a subset of the reference's
    order = np.argsort(ref.sp[0].weights)
    pos = np.searchsorted(ref.sp[0].weights[order], ws[:n])
    idx = order[pos]
    assert np.array_equal(ref.sp[0].weights[idx], ws[:n])
    assert_close(ps[:, :n], ref.sp[0].momenta[:, idx], tol, "momenta")
    assert_close(xs[:, :n], ref.sp[0].coords[:, idx], tol, "coords")
    cnt = torch.tensor([float(n)], device="cuda", dtype=torch.float64)
    dist.all_reduce(cnt)
    assert int(cnt.item()) == ref.sp[0].weights.size, (int(cnt.item()), ref.sp[0].weights.size)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("OK", name, "world", world, "slab", bool(slab), "window", window, backend)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 0, sys.argv[4] if len(sys.argv) > 4 else "nccl")
