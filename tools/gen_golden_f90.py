#!/usr/bin/env python
"""Golden vectors from the reference's OWN Fortran (build container only: needs /root/reference).

Every hot-path subroutine of /root/reference/f90/*.f90 is executed by oracle/f90py.py (a mechanical Fortran ->
Python/numpy translator; the image has no Fortran compiler) on the seeded inputs of tests/f90_cases.py, and a short
make_halfstep + make_step sequence of the reference's driver order (tests/pic_ref.py) is run on top of it.  Outputs go
to tests/golden/f90_kernels.npz and tests/golden/f90_steps.npz; tests/test_f90_golden.py checks the C++ oracle and
the CUDA library against them."""
import copy
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def step_cases():
    """(id, setup name, kwargs) of the step-level sequences"""
    return [("real_m2", "real_m2", dict(ions=True)), ("real_m3", "real_m3", dict(ions=False)),
            ("env_m1", "env_m1", dict(ions=False, undulator=True)), ("env_m3", "env_m3", dict(ions=False)),
            ("static_m2", "static_m2", dict(ions=False, px=50.0))]


def build_step_run(fim, sid, name, ions=False, undulator=False, px=0.0):
    from chimera_b200.solver_setup import SolverSetup
    from pic_ref import RefRun, RefSpecies
    from util import SETUPS, plasma, seed_fields

    S = SolverSetup(copy.deepcopy(SETUPS[name]))
    x, p, w = plasma(S, 1, 2, 11)
    keep = slice(None, None, 3)  # ~700 particles: every per-particle loop runs interpreted
    x, p, w = np.asfortranarray(x[:, keep]), np.asfortranarray(p[:, keep]), np.ascontiguousarray(w[keep])
    if name == "env_m1":
        p[0] += 391.0
    p[0] += px
    dev = (fim.undul_analytic, [0.3, 1.3, -1.0, 9.0]) if undulator else None
    sp = [RefSpecies(x, p, w, device=dev)]
    if ions:
        xi, pi_, wi = plasma(S, 1, 2, 18)
        sp.append(RefSpecies(np.asfortranarray(xi[:, keep]), 0 * np.asfortranarray(pi_[:, keep]), -np.ascontiguousarray(wi[keep]),
                             charge=1.0, mass=1886.0, still=True))
    run = RefRun(fim, S, sp, background=ions)
    if name != "static_m2":
        run.EG_fb[:] = seed_fields(S, 12, 0.5)
    return run, (px,) * len(sp)


def step_state(run):
    s = run.sp[0]
    return dict(J=run.J, Rho=run.Rho, EG_fb=run.EG_fb, EB=run.EB, coords=s.coords, coords_halfstep=s.coords_halfstep,
                momenta=s.momenta, weights=s.weights)


def main():
    from f90_cases import cases, fingerprint, flatten
    from oracle import fimera_f90

    F = fimera_f90.load()
    bad = {k: v for k, v in F._f90.warnings.items() if v}
    assert not bad, bad
    out, t0 = {}, time.time()
    for cid, fn, args in cases(F):
        res = getattr(F, fn)(*[a.copy(order="F") if isinstance(a, np.ndarray) else a for a in args])
        for k, arr in enumerate(flatten(res)):
            out["%s|%d" % (cid, k)] = fingerprint(arr)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "f90_kernels.npz"), **out)
    print("kernels: %d arrays, %.1f s" % (len(out), time.time() - t0))
    out, t0 = {}, time.time()
    for sid, name, kw in step_cases():
        run, px0 = build_step_run(F, sid, name, **kw)
        run.make_halfstep(px0=px0)
        for k, v in step_state(run).items():
            out["%s|half|%s" % (sid, k)] = fingerprint(v)
        for _ in range(2):
            run.make_step()
        for k, v in step_state(run).items():
            out["%s|step2|%s" % (sid, k)] = fingerprint(v)
        print(sid, "%.1f s" % (time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "f90_steps.npz"), **out)


if __name__ == "__main__":
    main()
