"""Shared helpers for the parity tests: seeded synthetic inputs with the edge cases the
reference's kernels special-case."""
import numpy as np

from chimera_b200.solver_setup import SolverSetup

TOL = 1e-12  # north_star: one-step parity within 1e-12 relative L2 in FP64


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    num = np.linalg.norm((a - b).ravel())
    return num / den if den > 0 else num


def assert_close(a, b, tol=TOL, what=""):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    e = rel_l2(a, b)
    assert e <= tol, "%s: rel L2 error %.3e > %.1e" % (what, e, tol)


SETUPS = {
    # real solver, 2 modes, space charge, chunked (LPA-like, lpa-testrun.py:30-33)
    "real_m2": dict(Grid=(-3.0, 1.0, 3.0, 0.05, 0.2), TimeStep=0.05, MaxAzimuthMode=1, Xchunked=(4, 3),
                    Features=("SpaceCharge", "StillAsBackground")),
    # real solver, 3 modes, no space charge
    "real_m3": dict(Grid=(-2.0, 1.0, 2.4, 0.06, 0.2), TimeStep=0.05, MaxAzimuthMode=2, Features=()),
    # real solver, single mode (Q7 territory)
    "real_m1": dict(Grid=(-2.0, 1.0, 2.4, 0.06, 0.2), TimeStep=0.05, MaxAzimuthMode=0, Features=("StaticKick",)),
    # envelope solver, Nko=0, Rcut (FEL-like, fel-testrun.py:35-41)
    "env_m1": dict(Grid=(-3.55, 3.55, 35.7, 7.1 / 40, 35.7 / 30), TimeStep=1.0 / 30, MaxAzimuthMode=0,
                   KxShift=2 * (391 / (1 + 1.95 ** 2 / 2) ** 0.5) ** 2, Rcut=25.0, CoPropagative=0.999,
                   Xchunked=(4, 6), Features={"NoPoissonCorrection": True}),
    # envelope solver with +-1 modes
    "env_m3": dict(Grid=(-2.0, 2.0, 6.0, 0.1, 0.3), TimeStep=0.05, MaxAzimuthMode=1, KxShift=30.0, Features=()),
}

_cache = {}


def setup(name):
    if name not in _cache:
        import copy

        _cache[name] = SolverSetup(copy.deepcopy(SETUPS[name]))
    return _cache[name]


def crandn(rng, shape):
    return np.asfortranarray(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))


def particles(S, n, seed, edge_cases=True, inside_only=False):
    """(coords(3,n), momenta(3,n), weights(n)) inside the solver grid, plus edge cases."""
    rng = np.random.default_rng(seed)
    a = S.Args
    nx, dx = a["Nx"], a["dx"]
    x = a["leftX"] + dx * (nx - 1) * rng.random(n)
    rmax = a["Rgrid"].max()
    r = rmax * np.sqrt(rng.random(n)) * 0.999
    th = 2 * np.pi * rng.random(n)
    coords = np.asfortranarray(np.vstack((x, r * np.cos(th), r * np.sin(th))))
    mom = np.asfortranarray(rng.standard_normal((3, n)) * 0.5)
    w = -1e-3 * (0.5 + rng.random(n))
    if edge_cases and n >= 16:
        coords[1:, 0] = 0.0                      # r = 0 exactly (phase special case)
        w[1] = 0.0                               # zero weight: skipped
        coords[1, 2], coords[2, 2] = rmax, 0.0   # r == Rgrid(nr): skipped
        coords[1, 3], coords[2, 3] = rmax * 2, 0.0
        mom[:, 4] = 0.0                          # zero momentum: skipped by dep_curr only
        coords[0, 5] = a["leftX"] + 0.3 * dx     # first cell
        coords[0, 6] = a["leftX"] + dx * (nx - 1) - 0.3 * dx  # last cell
        coords[1, 7], coords[2, 7] = 0.1 * a["dr"], 0.0       # innermost cell (ghost row contribution)
        coords[0, 8] = a["leftX"]                # exactly on a node
        if not inside_only:
            coords[0, 9] = a["leftX"] - 2.5 * dx   # outside (left): dropped
            coords[0, 10] = a["leftX"] + dx * (nx + 1.5)  # outside (right): dropped
    return coords, mom, np.asfortranarray(w)


def chunk_sorted(S, coords, mom, w, fim, nchnk):
    """Sort particles into x-chunks the way Specie.chunk_and_damp does (species.py:351-398)."""
    a = S.Args
    dom = np.asfortranarray([a["leftX"], a["rightX"], 0.0, a["Rgrid"].max() ** 2])
    ids, chunks, go_out = fim.chunk_coords_boundaries(coords, dom, a["Xgrid"], nchnk)
    order = np.argsort(ids, kind="stable")[go_out:]
    return (np.asfortranarray(coords[:, order]), np.asfortranarray(mom[:, order]), np.asfortranarray(w[order]),
            chunks)
