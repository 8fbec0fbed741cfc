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


def carrier_tol(S, tol=TOL):
    """Tolerance for quantities that went through the envelope gather/deposit.  Those multiply by the
    carrier exp(+-i kx0 x) (grid_deps_env.f90:47,219) with kx0 x up to ~4e5 rad in the FEL setups: a 1-ulp
    difference in x (any two compilers: FMA contraction, -ffast-math of the reference Makefile:14) moves
    the phase by eps * kx0 * |x|, so two correct implementations agree to that, not to 1e-12."""
    if not S.env:
        return tol
    a = S.Args
    return max(tol, np.finfo(float).eps * abs(a["kx0"]) * max(abs(a["leftX"]), abs(a["rightX"])))


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
    # quasi-static field of a relativistic beam (space-charge demo, first stage: 'StaticKick'), chunked
    "static_m2": dict(Grid=(-3.0, 1.0, 3.0, 0.05, 0.2), TimeStep=0.05, MaxAzimuthMode=1, Xchunked=(4, 3),
                      Features=("StaticKick",)),
    # envelope solver with +-1 modes
    "env_m3": dict(Grid=(-2.0, 2.0, 6.0, 0.1, 0.3), TimeStep=0.05, MaxAzimuthMode=1, KxShift=30.0, Features=()),
}

ALL_SETUPS = list(SETUPS)
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


# ---- inputs of the step-level (engine / driver) parity tests --------------------------------------
def plasma(S, ppc_x, ppc_r, seed, frac=(0.15, 0.85), thermal=0.05, dens=0.005):
    """Uniform plasma slab with unique (label) weights; coordinates inside the grid."""
    rng = np.random.default_rng(seed)
    a = S.Args
    nx, nr, dx, dr = a["Nx"], a["Nr"], a["dx"], a["dr"]
    ix = np.arange(int(frac[0] * nx), int(frac[1] * nx))
    ir = np.arange(0, int(0.8 * (nr - 1)))
    X, R, px, pr = np.meshgrid(ix, ir, np.arange(ppc_x), np.arange(ppc_r), indexing="ij")
    x = a["leftX"] + dx * (X + (px + 0.5) / ppc_x)
    r = dr * (R + (pr + 0.5) / ppc_r)
    x, r = x.ravel(), r.ravel()
    # jitter: on a perfectly regular lattice the envelope deposit (carrier exp(-i kx0 x)) cancels almost
    # exactly and the relative error of the tiny remainder is meaningless
    x = x + dx * 0.4 / ppc_x * (rng.random(x.size) - 0.5)
    r = r + dr * 0.4 / ppc_r * (rng.random(x.size) - 0.5)
    th = 2 * np.pi * rng.random(x.size)
    coords = np.asfortranarray(np.vstack((x, r * np.cos(th), r * np.sin(th))))
    mom = np.asfortranarray(thermal * rng.standard_normal((3, x.size)))
    w = -dens * dr * dx * 2 * np.pi * r / (ppc_x * ppc_r) * (1.0 + 1e-3 * rng.random(x.size))
    return coords, mom, np.asfortranarray(w)


def seed_fields(S, seed, amp=0.5):
    """A smooth, band-limited initial EG_fb so that gather/push see non-trivial fields."""
    rng = np.random.default_rng(seed)
    nx, nkr, nm = S.shape_fb
    eg = S.zeros_fb(6)
    kx = np.fft.fftfreq(nx) * nx
    env = np.exp(-(kx / (0.08 * nx)) ** 2)[:, None, None, None] * np.exp(-(np.arange(nkr) / (0.2 * nkr)) ** 2)[None, :, None, None]
    eg[:] = amp * env * (rng.standard_normal(eg.shape) + 1j * rng.standard_normal(eg.shape))
    return np.asfortranarray(eg)


def match(w_ref, w_eng):
    """permutation taking engine order to reference order via the (unique) weights"""
    a, b = np.argsort(w_ref, kind="stable"), np.argsort(w_eng, kind="stable")
    perm = np.empty_like(a)
    perm[a] = b
    assert np.array_equal(w_ref, w_eng[perm])
    return perm


def device_cases(rng, n=4000):
    """(name, args after (coord, fld, t)) for every routine of devices.f90 on a beam spread over the device entry,
    body and exit; shared by the CPU cross-check, the GPU parity test and the engine test."""
    x = np.asfortranarray(np.vstack((rng.random(n) * 16 - 3, rng.standard_normal(n) * 0.1, rng.standard_normal(n) * 0.1)))
    f = np.asfortranarray(rng.standard_normal((6, n)))
    nx = 96
    amap = np.asfortranarray(np.vstack((np.sin(0.7 * np.arange(nx)), np.cos(0.7 * np.arange(nx)))) * 1.3)
    cases = [
        ("undul_analytic", (np.array([1.95, 1.0, 1.0, 10.0]),)),
        ("undul_analytic_taper", (np.array([1.95, 1.0, 1.0, 10.0, 0.08]),)),
        ("undul_mapped", (amap, np.array([1.0, -1.0, 0.125]))),
        ("undul_mapped_tap", (amap, np.array([1.0, -1.0, 0.125, 9.0, -0.05]))),
        ("planewave", (np.array([0.8, 0.9, 0.5, 11.0, 2.0, 0.3, 0.4]),)),
        ("gaussbeam", (0.6, np.array([0.8, -1.0, 9.0, 0.02, -0.03, 3.0, 0.4, 0.5]))),
    ]
    return x, f, cases
