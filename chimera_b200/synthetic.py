"""Synthetic workloads of BASELINE.json: seeded plasmas, beams and laser seeds.

Everything here is input generation for benchmarks and tests; the arrays follow the reference's
conventions (``Specie.gen_parts`` species.py:132-215 for the plasma, ``Solver.add_gauss_beam``
solvers.py:555-603 for the laser) so that the reference driver could consume them unchanged.
"""
from __future__ import annotations

import numpy as np

# BASELINE.json configs[2]: "LWFA synthetic: Nz=4096, Nr=512, 3 azimuthal modes, 16 ppc (~1e8 particles)"
LWFA = dict(Nx=4096, Nr=512, modes=3, dx=0.04, dr=0.25, chunks=16, guards=10, density=0.005,
            a0=3.0, laser_Lx=4.0, laser_LR=16.0)


def lwfa_solver_config(nx=LWFA["Nx"], nr=LWFA["Nr"], modes=LWFA["modes"], dx=LWFA["dx"], dr=LWFA["dr"],
                       chunks=LWFA["chunks"], guards=LWFA["guards"]):
    """Solver dictionary (solvers.py:43-66 keys) of the LWFA synthetic case: real solver with space
    charge, still ions as background, x-chunked deposition, dt = dx (doc/tests/lpa-testrun.py:12-33)."""
    return {
        "Grid": (-nx * dx, 0.0, nr * dr, dx, dr), "TimeStep": dx, "MaxAzimuthMode": modes - 1,
        "Xchunked": (chunks, guards), "Features": ("SpaceCharge", "StillAsBackground"),
    }


def plasma_fixed_cell(args, cell=(2, 2, 4), density=LWFA["density"], thermal=0.05, seed=20260101,
                      margin_x=(12, 12), margin_r=8, x_cells=None, xp=np):
    """Uniform plasma, ``FixedCell=cell`` macro-particles per (x, r) cell with a random azimuth per cell
    (genparts, particle_tools.f90:84-128) and weights ``wght0 * r`` (species.py:118-121, 187).

    ``xp`` is numpy or torch-on-cuda; returns (coords(3,N), momenta(3,N), weights(N)) as arrays of
    the same library in the reference's Fortran (3,N) order, i.e. (N,3) row-major.
    """
    nx, nr, dx, dr, left = args["Nx"], args["Nr"], args["dx"], args["dr"], args["leftX"]
    nxp, nrp, nop = cell
    ix0, ix1 = margin_x[0], nx - 1 - margin_x[1]
    if x_cells is not None:
        ix0, ix1 = x_cells
    ir1 = nr - 1 - margin_r
    is_torch = xp.__name__ == "torch"
    if is_torch:
        import torch

        dev = torch.device("cuda")
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        ar = lambda n: torch.arange(n, device=dev, dtype=torch.float64)  # noqa: E731
        rand = lambda *s: torch.rand(*s, device=dev, dtype=torch.float64, generator=g)  # noqa: E731
        randn = lambda *s: torch.randn(*s, device=dev, dtype=torch.float64, generator=g)  # noqa: E731
    else:
        rng = np.random.default_rng(seed)
        ar = lambda n: np.arange(n, dtype=np.float64)  # noqa: E731
        rand = lambda *s: rng.random(s)  # noqa: E731
        randn = lambda *s: rng.standard_normal(s)  # noqa: E731
    ncx, ncr = ix1 - ix0, ir1
    cx = (ix0 + ar(ncx))[:, None, None, None, None]
    cr = ar(ncr)[None, :, None, None, None]
    px = ((ar(nxp) + 0.5) / nxp)[None, None, :, None, None]
    pr = ((ar(nrp) + 0.5) / nrp)[None, None, None, :, None]
    po = (ar(nop) / nop)[None, None, None, None, :]
    x = left + dx * (cx + px)
    r = dr * (cr + pr)                       # Rgrid(ir) + dr/2 + dr*packR, Rgrid(ir) = dr (ir - 1/2)
    th = 2 * np.pi * (rand(ncx, ncr)[:, :, None, None, None] + po)
    shape = (ncx, ncr, nxp, nrp, nop)
    if is_torch:
        import torch

        x, r, th = (t.expand(shape).reshape(-1) for t in (x, r, th))
        coords = torch.stack((x, r * torch.sin(th), r * torch.cos(th)), dim=1).contiguous()
        n = coords.shape[0]
        mom = (thermal * randn(n, 3)).contiguous()
    else:
        x, r, th = (np.broadcast_to(t, shape).reshape(-1) for t in (x, r, th))
        coords = np.ascontiguousarray(np.stack((x, r * np.sin(th), r * np.cos(th)), axis=1))
        n = coords.shape[0]
        mom = thermal * randn(n, 3)
    wght0 = -1.0 * density * dr * dx * 2 * np.pi / (nxp * nrp * nop)   # Charge = -1 (species.py:118)
    w = wght0 * r
    if not is_torch:
        return np.asfortranarray(coords.T), np.asfortranarray(mom.T), np.ascontiguousarray(w)
    return coords, mom, w.contiguous()


def laser_seed(setup, fim, a0=LWFA["a0"], k0=1.0, x0=None, Lx=LWFA["laser_Lx"], LR=LWFA["laser_LR"]):
    """Forward-propagating Gaussian pulse in EG_fb, following ``Solver.add_gauss_beam`` for the real
    solver (solvers.py:586-603) without its divergence cleaning: the pulse is the z component of
    mode 0.  ``fim`` supplies ``fb_scl_in`` (the CUDA drop-in, or the oracle in tests)."""
    a = setup.Args
    xg, rg, nx = a["Xgrid"], a["Rgrid"], a["Nx"]
    if x0 is None:
        x0 = xg[-1] - 4.0 * Lx
    k0w, a0w = 2 * np.pi * k0, 2 * np.pi * a0
    scl = setup.zeros_sp()
    dxx = xg[:, None] - x0
    scl[:, :, 0] = (a0w * np.cos(k0w * dxx) * np.exp(-dxx ** 2 / Lx ** 2 - rg[None, :] ** 2 / LR ** 2)
                    * (np.abs(rg[None, :]) < 3.5 * LR) * (np.abs(dxx) < 3.5 * Lx))
    scl[:, 0, 0] = 0.0
    scl_fb = fim.fb_scl_in(setup.zeros_fb(), scl, a["leftX"], *a["FBIn"])
    eg = setup.zeros_fb(6)
    eg[..., 2] = scl_fb / nx
    kxg = a["kx_g"][:, :, None]
    dt_op = -1j * a["w"] * np.sign(kxg + (kxg == 0))
    eg[..., 5] = dt_op * eg[..., 2]
    return eg


# ---------------------------------------------------------------------------------------------------------------
# The BASELINE.json configurations at their OWN grid shapes.  Solver dictionaries are the ones the reference's
# scripts build (cited per builder); particles are seeded synthetic plasmas / beams of the named shapes and sizes (the
# reference's own generators use the unseeded global numpy RNG).  Each builder returns a dict:
#     cfg      solver dictionary (solvers.py:43-66 keys)
#     species  list of dicts(coords, momenta, weights, charge, mass, still), or None with a 'beam' / 'plasma' recipe
#     laser    add_gauss_beam dictionary or None (solvers.py:555-603)
#     window   (velocity, staged) of a frame that moves every step, or None (chimera_main.py:40-51)
#     device   ('undul_analytic', [K0, lam_u, X0, Periods]) or None
#     px0      MomentaMeans[0] per species (make_halfstep's static kick, chimera_main.py:73-75)
# Used by bench.py --config and by tests/test_gpu_baseline_shapes.py.
# ---------------------------------------------------------------------------------------------------------------
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
    return dict(cfg=lwfa_solver_config(), species=None, laser={"a0": 3.0, "k0": 1.0, "x0": -20.0, "x_foc": 0.0, "Lx": 4.0, "LR": 16.0},
                window=None, device=None, px0=(0.0,), plasma=dict(cell=ppc_cell, density=0.005, ions=False))


def baseline_case(name):
    """builder by name: c1a (FEL stage), c1b (LPA stage), c2_static / c2_pic (space-charge drift), c3 (LWFA synthetic)"""
    return {"c1a": c1a_fel, "c1b": c1b_lpa, "c2_static": lambda: c2_space_charge("static"),
            "c2_pic": lambda: c2_space_charge("pic"), "c3": c3_lwfa}[name]()


def baseline_species(setup, case, seed=7):
    """the particle species of a baseline case on a SolverSetup"""
    if case["species"] is not None:
        return case["species"]
    if "beam" in case:
        return gaussian_beam(setup, seed=seed, **case["beam"])
    p = case["plasma"]
    return fill_plasma(setup, p["cell"], p["density"], seed, ions=p.get("ions", True))
