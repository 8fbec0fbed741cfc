"""Seeded inputs for every hot-path subroutine of the reference's f90 files (TEST INFRASTRUCTURE).

`cases()` yields (case id, fimera function name, argument tuple).  tools/gen_golden_f90.py feeds them to the
reference's Fortran executed by oracle/f90py.py and stores the OUTPUTS in tests/golden/f90_kernels.npz;
tests/test_f90_golden.py rebuilds the same inputs from the seeds and checks the C++ oracle (CPU) and the CUDA
library (GPU) against the stored outputs.  Inputs stay inside what the reference defines: no particle outside the
grid for the unguarded deposits (the Fortran writes out of bounds there), no Nko = 0 for the real-solver fb_rot /
fb_grad / fb_div (Q7: they read mode slot 1 out of bounds)."""
import numpy as np

from util import chunk_sorted, crandn, device_cases, particles, setup

SETUPS = ["real_m2", "real_m3", "env_m1", "env_m3"]


def cases(ofim):
    """ofim: any fimera backend for the chunk-sorting helper (index arithmetic only, bit-exact in all backends)"""
    rng = np.random.default_rng(4242)
    n = 240
    p3 = np.asfortranarray(rng.standard_normal((3, n)) * 3)
    f6 = np.asfortranarray(rng.standard_normal((6, n)) * 2)
    x3 = np.asfortranarray(rng.standard_normal((3, n)))
    yield "push_velocs", "push_velocs", (p3, f6, 0.37)
    yield "push_coords", "push_coords", (x3, p3, np.zeros((3, n), order="F"), 0.05)
    S = setup("real_m2")
    a = S.Args
    x, p, w = particles(S, 400, 5)
    dom = np.asfortranarray([a["leftX"] + 0.5, a["rightX"] - 0.3, 0.0, (0.8 * a["Rgrid"].max()) ** 2])
    yield "sortpartsout", "sortpartsout", (x, dom)
    wz = w.copy()
    wz[::7] = 0.0
    yield "sortoutghosts", "sortoutghosts", (wz,)
    dom = np.asfortranarray([a["leftX"], a["rightX"], 0.0, a["Rgrid"].max() ** 2])
    for nchnk in (1, 4):
        yield "chunk_coords_boundaries_%d" % nchnk, "chunk_coords_boundaries", (x, dom, a["Xgrid"], nchnk)
    idx = rng.permutation(400)[:350].astype(np.int64)
    yield "align_data_vec", "align_data_vec", (np.asfortranarray(rng.standard_normal((3, 400))), idx)
    yield "align_data_scl", "align_data_scl", (np.asfortranarray(rng.standard_normal(400)), idx)
    xg, rg_ = a["Xgrid"][:10], a["Rgrid"][:6]
    px, pr, po = np.mgrid[1:2:2j, 1:2:2j, 1:4:4j]
    yield "genparts", "genparts", (np.zeros((4, 10 * 6 * 16), order="F"), xg, rg_, np.asfortranarray(rng.random((10, 6))),
                                   np.asfortranarray((px.ravel() - 0.5) / 2), np.asfortranarray((pr.ravel() - 0.5) / 2),
                                   np.asfortranarray(np.exp(2j * np.pi * (po.ravel() - 1) / 4)))
    for name in SETUPS:
        S = setup(name)
        a = S.Args
        sfx = "_env" if S.env else ""
        r = np.random.default_rng(abs(hash(name)) % 1000 if False else sum(map(ord, name)))
        x, p, w = particles(S, 300, 7, inside_only=True)
        J0 = crandn(r, S.shape_sp + (3,)) * 1e-4
        R0 = crandn(r, S.shape_sp) * 1e-4
        yield name + ":dep_curr", "dep_curr" + sfx, (x, p, w, J0, a["leftX"], *a["DepProj"])
        yield name + ":dep_dens", "dep_dens" + sfx, (x, w, R0, a["leftX"], *a["DepProj"])
        for guards in ((0, 3) if a["Nx"] % 4 == 0 else ()):  # Nx not a multiple of nchnk: the reference indexes out of bounds
            xs, ps, ws, ch = chunk_sorted(S, *particles(S, 300, 10, inside_only=True), ofim, 4)
            xs[0] += a["dx"] * guards * (np.random.default_rng(11).random(xs.shape[1]) - 0.5) * 1.9
            yield name + ":dep_curr_chnk_g%d" % guards, "dep_curr" + sfx + "_chnk", (xs, ps, ws, S.zeros_sp(3), ch, guards, a["leftX"], *a["DepProj"])
            yield name + ":dep_dens_chnk_g%d" % guards, "dep_dens" + sfx + "_chnk", (xs, ws, S.zeros_sp(), ch, guards, a["leftX"], *a["DepProj"])
        Fd = crandn(r, S.shape_sp + (6,))
        xg_, _, wg_ = particles(S, 300, 12)  # with the edge cases: the gather guards them
        keep = (xg_[0] >= a["leftX"]) & (xg_[0] < a["leftX"] + a["dx"] * (a["Nx"] - 1))
        xg_, wg_ = np.asfortranarray(xg_[:, keep]), np.ascontiguousarray(wg_[keep])
        yield name + ":proj_fld", "proj_fld" + sfx, (xg_, wg_, Fd, np.asfortranarray(r.standard_normal((6, wg_.size))), a["leftX"], *a["DepProj"])
        yield name + ":eb_correction", "eb_correction" + sfx, (crandn(r, S.shape_sp + (6,)),)
        V, Sc = crandn(r, S.shape_fb + (3,)), crandn(r, S.shape_fb)
        yield name + ":fb_rot", "fb_rot" + sfx, (S.zeros_fb(3), V, *a["FBDiff"])
        yield name + ":fb_grad", "fb_grad" + sfx, (S.zeros_fb(3), Sc, *a["FBDiff"])
        yield name + ":fb_div", "fb_div" + sfx, (S.zeros_fb(), V, *a["FBDiff"])
        yield name + ":fb_graddiv", "fb_graddiv" + sfx, (V, *a["FBDiff"])
        Jg, Rg = crandn(r, S.shape_sp + (3,)), crandn(r, S.shape_sp)
        yield name + ":fb_vec_in", "fb_vec_in", (S.zeros_fb(3), Jg, a["leftX"], *a["FBCurrIn"])
        yield name + ":fb_scl_in", "fb_scl_in", (S.zeros_fb(), Rg, a["leftX"], *a["FBIn"])
        EG, B = crandn(r, S.shape_fb + (6,)), crandn(r, S.shape_fb + (3,))
        yield name + ":fb_vec_out", "fb_vec_out", (V, a["leftX"], *a["FBout"])
        yield name + ":fb_scl_out", "fb_scl_out", (Sc, a["leftX"], *a["FBout"])
        yield name + ":fb_eb_out", "fb_eb_out", (S.zeros_sp(6), EG, B, a["leftX"], *a["FBout"])
        nf = 12
        g = np.arange(nf)
        filt = (g >= 0.75 * nf) * (0.5 - 0.5 * np.cos(np.pi * (g - 0.75 * nf) / (0.25 * nf))) ** 2
        yield name + ":fb_filtr", "fb_filtr", (V, a["leftX"], a["kx"], filt, 0)  # mode 0 is the one the driver uses
        g1, g2 = crandn(r, S.shape_fb + (3,)), crandn(r, S.shape_fb + (3,))
        if S.space_charge:
            yield name + ":maxwell_push_with_spchrg", "maxwell_push_with_spchrg", (EG, V, g1, g2, S.PSATD_E, S.PSATD_G)
        else:
            yield name + ":maxwell_push_wo_spchrg", "maxwell_push_wo_spchrg", (EG, V, S.PSATD_E, S.PSATD_G)
        c1, c2 = S.static_coeffs(50.0)
        yield name + ":maxwell_init_push", "maxwell_init_push", (EG, V, g1, c1, c2)
        yield name + ":poiss_corr", "poiss_corr", (V, g1, g2, EG[..., :3], a["dt_inv"], a["PoissFact"])
        yield name + ":poiss_corr_stat", "poiss_corr_stat", (V, g1, g2, -1j * 0.9998 * a["kx"], a["PoissFact"])
        yield name + ":field_drift", "field_drift", (EG, a["kx"], 0.9998, a["TimeStep"])
        yield name + ":omp_mult_vec", "omp_mult_vec", (V, a["DepFact"])
        yield name + ":omp_mult_scl", "omp_mult_scl", (Sc, a["DepFact"])
        yield name + ":omp_add_vec", "omp_add_vec", (V, g1)
        yield name + ":omp_add_scl", "omp_add_scl", (Sc, g1[..., 2])
    xd, fd, dev = device_cases(np.random.default_rng(16), n=300)
    for name, args in dev:
        yield "device:" + name, name, (xd, fd, 0.3, *args)
    # utils.f90 diagnostics helpers and SR.f90 (NEXT rows 3 and 4)
    r = np.random.default_rng(55)
    for nm in (1, 3):
        fld = np.asfortranarray(r.standard_normal((17, 9, nm, 3)) + 1j * r.standard_normal((17, 9, nm, 3)))
        yield "intens_profo_m%d" % nm, "intens_profo", (fld, 12)
    yield "density_2x", "density_2x", (r.uniform(-1.2, 1.2, 900), r.uniform(-0.7, 2.4, 900), r.random(900),
                                       np.array([-1.0, 1.0, -0.5, 2.0]), 20, 13)
    from test_sr import far_grid, near_args, tracks

    x, mp, mn, w, dt = tracks(40, 3, 43)
    g = far_grid(7, 2, 3)
    s0 = np.asfortranarray(r.random((7, 2, 3)))
    yield "sr_far_tot", "sr_calc_far_tot", (s0, x, mp, mn, w, dt, *g)
    yield "sr_far_comp2", "sr_calc_far_comp", (s0, x, mp, mn, w, 2, dt, *g)
    for circ in (False, True):
        na = near_args(circ, 6, 3, 2)
        s1 = np.asfortranarray(r.random((6, 3, 2)))
        tag = "nearcirc" if circ else "near"
        yield "sr_%s_tot" % tag, "sr_calc_%s_tot" % tag, (s1, x, mn, w, dt, *na)
        yield "sr_%s_comp3" % tag, "sr_calc_%s_comp" % tag, (s1, x, mn, w, 3, dt, *na)


def flatten(out):
    """outputs of one call as a list of arrays (tuple / scalar / single array)"""
    if isinstance(out, tuple):
        return [np.asarray(o) for o in out]
    return [np.asarray(out)]


# ---- compact fixtures: small outputs are stored whole, large ones as 64 seeded random projections + the L2 norm.
# For candidate = reference + e the projection differences are N(0, |e|^2), so rms(diff) / |reference| estimates the
# relative L2 error (to ~12 %) without keeping megabytes of incompressible doubles in git.
NPROJ, WHOLE = 64, 2048


def _proj(a):
    v = np.asarray(a, dtype=np.complex128 if np.iscomplexobj(a) else np.float64).ravel(order="F")
    rng = np.random.default_rng(v.size)
    out = np.zeros(NPROJ, dtype=v.dtype)
    for k in range(NPROJ):
        out[k] = rng.standard_normal(v.size) @ v
    return out


def fingerprint(a):
    a = np.asarray(a)
    if a.size <= WHOLE:
        return a.copy()
    return np.concatenate(([np.linalg.norm(a.ravel())], _proj(a)))


def fingerprint_error(got, stored):
    """relative L2 error of `got` against what `fingerprint(reference)` stored"""
    got = np.asarray(got)
    if got.size <= WHOLE:
        assert stored.shape == got.shape, (stored.shape, got.shape)
        if got.dtype.kind in "iub":
            return 0.0 if np.array_equal(got, stored) else 1.0
        den = np.linalg.norm(stored.ravel())
        num = np.linalg.norm((got - stored).ravel())
        return num / den if den > 0 else num
    assert stored.shape == (NPROJ + 1,), (stored.shape, got.shape)
    norm, ref = abs(stored[0]), stored[1:]
    d = _proj(got) - ref
    return float(np.sqrt(np.mean(np.abs(d) ** 2)) / (norm if norm > 0 else 1.0))
