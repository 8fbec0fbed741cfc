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
