"""The BASELINE.json configurations at their OWN grid shapes (TEST INFRASTRUCTURE).

Solver dictionaries are the ones the reference's scripts build (cited per builder); particles are seeded
synthetic plasmas / beams of the named shapes and sizes (the reference's own generators use the unseeded global
numpy RNG and the reference checkout does not exist on the GPU box).  Each builder returns a dict:

    cfg      solver dictionary (solvers.py:43-66 keys)
    species  list of dicts(coords, momenta, weights, charge, mass, still)
    laser    add_gauss_beam dictionary or None (solvers.py:555-603)
    window   (velocity, staged) of a frame that moves every step, or None (chimera_main.py:40-51)
    device   ('undul_analytic', [K0, lam_u, X0, Periods]) or None
    px0      MomentaMeans[0] per species (make_halfstep's static kick, chimera_main.py:73-75)
"""
import numpy as np


def _cell_plasma(leftX, dx, dr, ix0, ix1, ir1, cell, density, charge, thermal, rng, profile=None):
    """FixedCell plasma (species.py:95-107, particle_tools.f90:84-128): cell = (nx_p, nr_p, no_p) particles per
    (x, r) cell on a regular sub-lattice, one random azimuth offset per cell, weight = wght0 * r."""
    nxp, nrp, nop = cell
    cx = np.arange(ix0, ix1)[:, None, None, None, None]
    cr = np.arange(0, ir1)[None, :, None, None, None]
    px = ((np.arange(nxp) + 0.5) / nxp)[None, None, :, None, None]
    pr = ((np.arange(nrp) + 0.5) / nrp)[None, None, None, :, None]
    po = (np.arange(nop) / nop)[None, None, None, None, :]
    shape = (ix1 - ix0, ir1, nxp, nrp, nop)
    x = np.broadcast_to(leftX + dx * (cx + px), shape).ravel()
    r = np.broadcast_to(dr * (cr + pr), shape).ravel()
    th = np.broadcast_to(2 * np.pi * (rng.random((ix1 - ix0, ir1))[:, :, None, None, None] + po), shape).ravel()
    coords = np.asfortranarray(np.vstack((x, r * np.sin(th), r * np.cos(th))))
    n = x.size
    mom = np.asfortranarray(thermal * rng.standard_normal((3, n)))
    w = charge * density * dr * dx * 2 * np.pi / (nxp * nrp * nop) * r
    if profile is not None:
        w = w * profile(coords[0], coords[1], coords[2])
    # unique weights: the tests match engine and reference particle order through them
    w = w * (1.0 + 1e-9 * rng.random(n))
    keep = w != 0
    return np.asfortranarray(coords[:, keep]), np.asfortranarray(mom[:, keep]), np.ascontiguousarray(w[keep])


def c1a_fel(periods=10):
    """BASELINE configs[0], FEL stage: doc/tests/fel-testrun.py:12-63 (= doc/fel-lpa-demo.ipynb cells 12-19):
    envelope solver Nx=120, Nr=120 cut to Rg_cut, one mode, co-propagative, no Poisson correction, analytic
    undulator, 'Staged' frame every step; beam of 2 x 48 x 120 x 50 = 5.8e5 macro-particles ('RandCell': 50,
    doubled by denoise)."""
    K0, lam0 = 1.95, 2.8
    g0 = 200 / 0.511
    lbx = lbr = 80e-4 / lam0
    dens = 20e-12 / 1.6022e-19 / (np.pi * 80e-4 ** 3) / (1.1e21 / 2.8e4 ** 2)
    gg = g0 / (1.0 + K0 ** 2 / 2) ** 0.5
    k_res = 2 * gg ** 2
    vb = (1.0 - gg ** -2) ** 0.5
    Lgx, Rg, Rg_cut = 200e-4 / lam0, 1000e-4 / lam0, 700e-4 / lam0
    Nx = Nr = 120
    dt = 1.0 / 30
    cfg = {"Grid": (-0.5 * Lgx, 0.5 * Lgx, Rg, Lgx / Nx, Rg / Nr), "TimeStep": dt, "MaxAzimuthMode": 0,
           "KxShift": k_res, "Rcut": Rg_cut, "CoPropagative": vb, "Xchunked": (4, 6),
           "Features": {"NoPoissonCorrection": True}}
    seed = {"a0": 0.15, "k0": k_res, "x0": -30e-4 / lam0, "x_foc": 70.0 / lam0, "Lx": 15e-4 / lam0, "LR": 180e-4 / lam0}
    rng = np.random.default_rng(20260101)
    # gen_randcell: 50 random particles per cell of the beam's own grid (dx = Lgx/Nx, dr = lbr/Nr) inside
    # |x| < lbx/2, r < lbr; weights = wght0 * r
    bdx, bdr = Lgx / Nx, lbr / Nr
    ncx = int(round(lbx / bdx))
    n = ncx * Nr * 50
    x = -0.5 * lbx + lbx * rng.random(n)
    r = lbr * rng.random(n)
    th = 2 * np.pi * rng.random(n)
    w = -1.0 * dens * bdr * bdx * 2 * np.pi / 50 * r * (1.0 + 1e-9 * rng.random(n))
    mom = np.vstack((g0 * (1 + 1e-4 * rng.standard_normal(n)), 2e-5 * g0 * rng.standard_normal(n),
                     2e-5 * g0 * rng.standard_normal(n)))
    # denoise((k_res,)) (species.py:322-349): a copy displaced by half the resonant wavelength with the same weight
    # cancels the shot noise at k_res
    y, z = r * np.sin(th), r * np.cos(th)
    coords = np.asfortranarray(np.vstack((np.r_[x, x + 0.5 / k_res], np.r_[y, y], np.r_[z, z])))
    mom = np.asfortranarray(np.hstack((mom, mom)))
    w = 0.5 * np.r_[w, w * (1 + 1e-10)]
    sp = dict(coords=coords, momenta=mom, weights=np.ascontiguousarray(w), charge=-1.0, mass=1.0, still=False)
    return dict(cfg=cfg, species=[sp], laser=seed, window=(vb, True), device=("undul_analytic", [K0, 1.0, 1.0, float(periods)]),
                px0=(g0,))


def c1b_lpa(nx_box=21.0):
    """BASELINE configs[0], LPA stage: doc/tests/lpa-testrun.py:12-62 (Nx=528, Nr=65, 2 modes; the notebook's box is
    longer, Nx=1272): real solver with SpaceCharge + StillAsBackground, Xchunked (4,10), a0=3 pulse, electrons
    'FixedCell': (2,2,4) = 16 per cell and equally many still ions.  The reference starts with empty species and lets
    the window feed plasma in; here the box is pre-filled so that all 100 test steps do real work, and the window
    (AbsorbLayer 175 cells, every 10 steps) feeds fresh cells on the right."""
    xgmin, xgmax, Rg, dx, dr = 1.0 - nx_box, 1.0, 16.0, 0.04, 0.25
    cfg = {"Grid": (xgmin, xgmax, Rg, dx, dr), "TimeStep": dx, "MaxAzimuthMode": 1, "Xchunked": (4, 10),
           "Features": ("SpaceCharge", "StillAsBackground")}
    laser = {"a0": 3.0, "k0": 1.0, "x0": -14.0, "x_foc": 45.0, "Lx": 4.0, "LR": 4.0}
    return dict(cfg=cfg, species=None, laser=laser, window=None, device=None, px0=(0.0, 0.0),
                plasma=dict(cell=(2, 2, 4), density=0.005, frame={"Steps": 10, "AbsorbLayer": 175}))


def fill_plasma(S, cell, density, seed, ix0=None, ix1=None, thermal=0.0, ions=True, margin_r=2):
    """electrons (+ still ions on top: 'IonsOnTop', chimera_main.py:262-270) in cells [ix0, ix1) of the solver grid"""
    a = S.Args
    rng = np.random.default_rng(seed)
    ix0 = 0 if ix0 is None else ix0
    ix1 = a["Nx"] - 1 if ix1 is None else ix1
    x, p, w = _cell_plasma(a["leftX"], a["dx"], a["dr"], ix0, ix1, a["Nr"] - 1 - margin_r, cell, density, -1.0, thermal, rng)
    out = [dict(coords=x, momenta=p, weights=w, charge=-1.0, mass=1.0, still=False)]
    if ions:
        out.append(dict(coords=x.copy(order="F"), momenta=np.zeros_like(p), weights=-w, charge=1.0, mass=1886.0, still=True))
    return out


def c2_space_charge(stage):
    """BASELINE configs[1]: doc/space-charge-demo(vs_ocelot).ipynb cell 7 (stage 'static': 'StaticKick', dt=1,
    21 steps) and cell 9 (stage 'pic': 'SpaceCharge', dt=0.06, 334 steps): Nx=304, Nr=301 nodes, 2 modes,
    Xchunked (4,6), 'Staged' frame every step at v=1, Gaussian beam sigma=3, px=50, 'FixedCell': (4,8,8) in
    |x| < 3.5 sigma, r < 3.5 sigma = 151 x 36 cells x 256 = 1.39e6 macro-particles."""
    Size, pz0 = 3.0, 50.0
    e = 1.602176634e-19
    nmax = 30e-12 / e / ((Size * 1e-4) ** 3 * (2 * np.pi) ** 1.5) / 1.1e21
    xmin, xmax, lrg = -7.0 * Size, 7.0 * Size, 30 * Size
    dx, dr = (xmax - xmin) / 300, lrg / 300
    dt = 1.0 if stage == "static" else 0.06
    cfg = {"Grid": (xmin, xmax, lrg, dx, dr), "TimeStep": dt, "MaxAzimuthMode": 1,
           "Features": ("StaticKick",) if stage == "static" else ("SpaceCharge",), "Xchunked": (4, 6)}
    return dict(cfg=cfg, species=None, laser=None, window=(1.0, True), device=None, px0=(pz0,),
                beam=dict(size=Size, px=pz0, density=nmax, cell=(4, 8, 8)))


def gaussian_beam(S, size, px, density, cell, seed):
    a = S.Args
    rng = np.random.default_rng(seed)
    xg = a["Xgrid"]
    ix0, ix1 = int((xg < -3.5 * size).sum()) - 1, int((xg < 3.5 * size).sum()) + 1
    ir1 = int((a["Rgrid"] < 3.5 * size).sum()) + 1
    prof = lambda x, y, z: np.exp(-0.5 * (x ** 2 + y ** 2 + z ** 2) / size ** 2)  # noqa: E731
    x, p, w = _cell_plasma(a["leftX"], a["dx"], a["dr"], ix0, ix1, ir1, cell, density, -1.0, 0.0, rng, profile=prof)
    p[0] += px
    return [dict(coords=x, momenta=p, weights=w, charge=-1.0, mass=1.0, still=False)]


def c3_lwfa(ppc_cell=(2, 2, 4)):
    """BASELINE configs[2]: LWFA synthetic, Nz=4096, Nr=512 (+ghost node = 513), 3 azimuthal modes, 16 per cell
    (3.3e7 macro-particles), real solver with SpaceCharge, Xchunked (16,10), dt = dx."""
    from chimera_b200.synthetic import lwfa_solver_config

    return dict(cfg=lwfa_solver_config(), species=None, laser={"a0": 3.0, "k0": 1.0, "x0": -20.0, "x_foc": 0.0, "Lx": 4.0, "LR": 16.0},
                window=None, device=None, px0=(0.0,), plasma=dict(cell=ppc_cell, density=0.005, ions=False))
