"""Host-side construction of everything the spectral kernels consume.

The reference builds these tables in ``Solver.__init__`` / ``get_spectral_operators`` /
``PSATD_coeffs`` (reference moduls/solvers.py:27-279, 660-759) with numpy/scipy and hands them to the
Fortran kernels positionally.  The GPU box has no reference checkout, so the benchmark, the
resident engine and the GPU tests need their own builder.  This is a from-scratch
implementation of the same mathematics, organised per azimuthal mode number rather than per
array slot; tests/test_solver_setup.py checks it slot-for-slot against the reference's
``Solver`` (in the build container) and against committed golden fixtures (everywhere).

Conventions: all arrays Fortran-ordered; radial node j sits at
r = dr (j - 1/2) with node 0 a ghost; spectral arrays are (Nx, Nkr, M); the operator stacks are
stored the way the kernels index them:  In/InCurr (Nr-1, Nkr, M), Out (Nkr, Nr-1, M),
DpS2S/DmS2S (Nkr, Nkr, M+1 | M+2).
"""
from __future__ import annotations

import numpy as np
from scipy.special import jn, jn_zeros

# CODATA values as in scipy.constants (the reference imports them, solvers.py:21)
_M_E = 9.1093837139e-31
_C = 299792458.0
_E = 1.602176634e-19
_EPS0 = 8.8541878188e-12
try:  # keep bit-identical to whatever scipy the reference would use here
    from scipy.constants import c as _C, e as _E, epsilon_0 as _EPS0, m_e as _M_E  # noqa: F811
except Exception:  # pragma: no cover
    pass


def mode_slots(nko: int, env: bool, ext: int = 0):
    """Azimuthal mode number held by each array slot (reference solvers.py:721-757).

    real solver: 0..nko+ext.  Envelope solver: -nko-ext..nko+ext, except that for nko == 0 the
    reference does not reorder the ext=1 stack, so its slots hold [0, +1, -1]
    (SURVEY.md Appendix A "Operators are consumed positionally").
    """
    if not env:
        return list(range(0, nko + ext + 1))
    lo, hi = -nko - ext, nko + ext
    if nko > 0:
        return list(range(lo, hi + 1))
    return [0] if ext == 0 else [0, 1, -1]


class SolverSetup:
    """Grid, spectral axes, DHT / mode-coupling operators and PSATD tables of one solver.

    Parameters follow the reference's solver dictionary (solvers.py:43-66): ``Grid`` =
    (leftX, rightX, lengthR, dx, dr), ``TimeStep``, ``MaxAzimuthMode``, optional ``Xchunked``,
    ``KxShift``, ``Rcut``, ``CoPropagative``, ``Features``.
    """

    def __init__(self, cfg: dict):
        a = self.Args = dict(cfg)
        feats = a.setdefault("Features", ())
        self.env = "KxShift" in a
        nko = a["Nko"] = int(a["MaxAzimuthMode"])
        leftX, rightX, lengthR, dx, dr = a["Grid"]
        dt = a["dt"] = a["TimeStep"]
        a["dx"], a["dr"] = dx, dr
        a["dx_inv"], a["dr_inv"], a["dt_inv"] = 1.0 / dx, 1.0 / dr, 1.0 / dt

        # x grid: even node count, divisible by the number of chunks (solvers.py:60-73)
        if "Xchunked" in a:
            nth = a["nthrds"] = a["Xchunked"][0]
            a["Nxchunk"] = 2 * int(np.round(0.5 / nth / dx * (rightX - leftX)))
            nx = a["Nxchunk"] * nth
        else:
            nx = int(2 * np.round(0.5 / dx * (rightX - leftX)))
        a["Nx"] = nx
        a["rightX"] = rightX
        a["Xgrid"] = rightX - dx * np.arange(nx)[::-1]
        a["leftX"] = a["Xgrid"][0]

        a["kx0"] = 2 * np.pi * a["KxShift"] if self.env else 0.0
        a["kx_env"] = 2 * np.pi * np.fft.fftfreq(nx, dx)
        a["kx"] = a["kx0"] + a["kx_env"]
        a["dkx"] = (a["kx"][1] - a["kx"][0]) / (2.0 * np.pi)

        # r grid with the half-cell offset (solvers.py:86-91)
        nkr = a["Nkr"] = int(np.round(lengthR / dr))
        rfull = a["RgridFull"] = dr * (np.arange(nkr + 1) - 0.5)
        lengthR = a["lengthR"] = rfull[-1] + dr

        ops1 = self._operators(ext=1)
        ops0 = self._operators(ext=0)
        a["DpS2S"], a["DmS2S"] = ops1["Dp"], ops1["Dm"]
        a["InFull"], a["OutFull"] = ops0["In"], ops0["Out"]
        a["w"], a["kr"], a["kr_g"], a["kx_g"], a["idxM"] = (ops0[k] for k in ("w", "kr", "kr_g", "kx_g", "idxM"))
        m_tot = a["Mtot"] = a["InFull"].shape[-1]
        a["PoissFact"] = 1.0 / a["w"] ** 2

        vgrid = 2 * np.pi * dx * dr * rfull
        a["VGrid"] = (rfull > 0.0) / vgrid
        a["InCurr"] = a["InFull"] * a["VGrid"][None, 1:, None]

        if "Rcut" in a:  # radial truncation of the real-space grid only (solvers.py:123-136)
            ncut = int((rfull < a["Rcut"]).sum())
            a["Rgrid"] = rfull[:ncut]
            a["In"] = a["InFull"][:, : ncut - 1]
            a["InCurr"] = a["InCurr"][:, : ncut - 1]
            a["Out"] = a["OutFull"][: ncut - 1]
            a["Rcut"] = a["Rgrid"].max()
        else:
            a["In"], a["Out"], a["Rgrid"] = a["InFull"], a["OutFull"], rfull
        a["Nr"] = a["Rgrid"].shape[0]
        a["lowerR"] = (a["Rgrid"] * (a["Rgrid"] >= 0)).min()
        a["upperR"] = a["Rgrid"].max()

        order = (np.abs(a["idxM"]) + 1)[None, None, :]
        a["EnergyFact"] = (
            0.5e-2 * (_M_E * _C ** 2 / _E) ** 2 * (4 * np.pi * _EPS0) * lengthR ** 2 / a["dkx"]
            * jn(order, a["kr_g"] * lengthR) ** 2
        )
        if not self.env:
            a["EnergyFact"] *= 0.5

        # kernels index (source, destination) -> swap the two leading axes (solvers.py:152-157)
        for k in ("DpS2S", "DmS2S", "InCurr", "In", "Out"):
            a[k] = np.swapaxes(a[k], 0, 1)
        for k, v in list(a.items()):
            if isinstance(v, np.ndarray):
                a[k] = np.asfortranarray(v)

        if self.env:
            filt_x = self._antialias()
            kx_base = a["kx_env"]
        else:
            filt_x = np.cos(0.5 * dx * a["kx_g"])[:, :, None] ** 2
            kx_base = a["kx"]
        a["DepFact"] = np.asfortranarray(
            (2 * np.pi) ** 2 / nx * filt_x * np.cos(0.5 * dr * a["kr_g"]) ** 2
        )
        a["FBDiff"] = [a["DpS2S"], a["DmS2S"], a["kx"]]
        a["DepProj"] = [a["Rgrid"], a["dx_inv"], a["dr_inv"]] + ([a["kx0"]] if self.env else [])
        a["FBCurrIn"] = (kx_base, a["InCurr"])
        a["FBIn"] = (kx_base, a["In"])
        a["FBout"] = (kx_base, a["Out"])
        # NB the reference never swaps the axes of OutFull (solvers.py:152 lists only In/Out/...), so
        # the diagnostics' FBoutFull pack carries it as (r, k, m); kept as is.
        a["FBoutFull"] = (kx_base, a["OutFull"])

        self.space_charge = "SpaceCharge" in feats
        self.PSATD_E, self.PSATD_G = self.psatd_coeffs(a.get("CoPropagative", 1.0))
        self.shape_sp = (nx, a["Nr"], m_tot)
        self.shape_fb = (nx, nkr, m_tot)

    # -- operators ---------------------------------------------------------------------------
    def _operators(self, ext):
        a = self.Args
        nx, nkr, nko, length_r, kx, rfull = (a[k] for k in ("Nx", "Nkr", "Nko", "lengthR", "kx", "RgridFull"))
        modes = mode_slots(nko, self.env, ext)
        r = rfull[1:, None]

        def zeros_of(m):  # radial wavenumbers of mode m: Bessel zeros / R (solvers.py:737)
            return jn_zeros(abs(int(m)), nkr) / length_r

        mtot = len(modes)
        out = np.zeros((nkr, nkr, mtot))
        inv_ = np.zeros_like(out)
        dp = np.zeros_like(out)
        dm = np.zeros_like(out)
        kr_g = np.zeros((nx, nkr, mtot))
        w = np.zeros((nx, nkr, mtot))
        for s, m in enumerate(modes):
            k_m = zeros_of(m)
            kr_g[:, :, s] = k_m[None, :]
            w[:, :, s] = np.sqrt(kx[:, None] ** 2 + k_m[None, :] ** 2)
            out[:, :, s] = jn(m, r * k_m[None, :])            # backward DHT  (r, k)
            inv_[:, :, s] = np.linalg.inv(out[:, :, s])       # forward DHT   (k, r)
            k_up, k_dn = zeros_of(m + 1), zeros_of(abs(m - 1))
            dp[:, :, s] = inv_[:, :, s].dot(0.5 * k_up[None, :] * jn(m, r * k_up[None, :]))
            dm[:, :, s] = inv_[:, :, s].dot(0.5 * k_dn[None, :] * jn(m, r * k_dn[None, :]))
        # `kr` of the reference carries one extra mode on each open side; keep its slot order
        if self.env:
            lo, hi = -nko - ext - 1, nko + ext + 1
            if nko > 0:
                kr_modes = list(range(lo + 1, hi))
            else:
                ncol = hi - lo + 1
                kr_modes = [(j if j <= hi else j - ncol) for j in range(ncol)]
        else:
            kr_modes = list(range(0, nko + ext + 2))
        kr = np.stack([zeros_of(m) for m in kr_modes], axis=1)
        kx_g = np.repeat(kx[:, None], nkr, axis=1)
        if self.env and nko > 0:
            idx = np.arange(-nko - ext, nko + ext + 1)
        elif self.env:
            idx = np.arange(-ext, ext + 1)
        else:
            idx = np.arange(0, nko + ext + 1)
        return dict(In=inv_, Out=out, Dp=dp, Dm=dm, w=w, kr=kr, kr_g=kr_g, kx_g=kx_g, idxM=idx)

    # -- PSATD coefficient tables (solvers.py:227-279) -----------------------------------------
    def psatd_coeffs(self, beta=1.0):
        a = self.Args
        w, dt = a["w"], a["TimeStep"]
        kxb = beta * a["kx_g"][:, :, None]
        s, c = np.sin(dt * w), np.cos(dt * w)
        if self.space_charge:
            ce = np.zeros(w.shape + (5,), dtype="double", order="F")
        elif self.env:
            ce = np.zeros(w.shape + (3,), dtype="complex", order="F")
        else:
            ce = np.zeros(w.shape + (3,), dtype="double", order="F")
        cg = np.zeros_like(ce)
        ce[..., 0], ce[..., 1] = c, s / w
        cg[..., 0], cg[..., 1] = -w * s, c
        if self.env:
            den = w ** 2 - kxb ** 2
            ph_h, ph = np.exp(-0.5j * kxb * dt), np.exp(1j * kxb * dt)
            ce[..., 2] = ph_h * 1j * kxb / den * (1.0 - ph * (w / (1j * kxb) * s + c))
            cg[..., 2] = ph_h * w ** 2 / den * (1.0 + ph * (1j * kxb / w * s - c))
        else:
            ce[..., 2] = -s / w
            cg[..., 2] = 1 - c
        if self.space_charge:
            ce[..., 3] = (dt * w * c - s) / w ** 3 / dt
            ce[..., 4] = (s - dt * w) / w ** 3 / dt
            cg[..., 3] = (1 - c - dt * w * s) / w ** 2 / dt
            cg[..., 4] = (c - 1) / w ** 2 / dt
        return ce, cg

    # -- laser injection (solvers.py:555-603) ----------------------------------------------------
    def add_gauss_beam(self, fim, laser, EG_fb=None):
        """``Solver.add_gauss_beam``: a Gaussian pulse ``laser = {'a0','k0','x0','x_foc','Lx','LR'}`` added to ``EG_fb``
        (a new array when none is given), propagated analytically from its focus to ``x0``.  ``fim`` is the fimera
        backend that supplies ``fb_scl_in``, ``fb_graddiv[_env]``, ``omp_mult_vec`` and ``omp_add_vec`` (the CUDA
        drop-in, or the oracle in tests) -- the same calls the reference makes, in the same order."""
        from scipy.special import j1

        a = self.Args
        k0, a0 = 2 * np.pi * laser["k0"], 2 * np.pi * laser["a0"]
        x_focus = laser["x0"] - laser["x_foc"]
        kx_g, kr_g, w = a["kx_g"], a["kr_g"], a["w"][:, :, :, None]
        vec_fb = self.zeros_fb(3)
        if self.env:
            nko = a["Nko"]
            e_s0 = a0 * 0.5 * np.pi ** 0.5 * laser["Lx"] * laser["LR"] ** 2 * a["dkx"] / a["lengthR"] ** 2
            vec_fb[:, :, nko, 2] = (e_s0 / j1(a["lengthR"] * kr_g[:, :, nko]) ** 2 * np.exp(-1j * kx_g * laser["x0"])
                                    * np.exp(-0.25 * (kx_g - k0) ** 2 * laser["Lx"] ** 2
                                             - 0.25 * kr_g[:, :, nko] ** 2 * laser["LR"] ** 2))
            dt_op = -1j * w
        else:
            xg, rg = a["Xgrid"], a["Rgrid"]
            scl = self.zeros_sp()
            dxx = xg[:, None] - laser["x0"]
            scl[:, :, 0] = (a0 * np.cos(k0 * dxx) * np.exp(-dxx ** 2 / laser["Lx"] ** 2 - rg[None, :] ** 2 / laser["LR"] ** 2)
                            * (np.abs(rg[None, :]) < 3.5 * laser["LR"]) * (np.abs(dxx) < 3.5 * laser["Lx"]))
            scl[:, 0, 0] = 0.0
            scl_fb = fim.fb_scl_in(self.zeros_fb(), scl, a["leftX"], *a["FBIn"])
            vec_fb[:, :, :, 2] = scl_fb / a["Nx"]
            kxg4 = kx_g[:, :, None, None]
            dt_op = -1j * w * np.sign(kxg4 + (kxg4 == 0))
        ee = vec_fb.copy(order="F")
        # div_clean (solvers.py:644-650): E += PoissFact * grad div E
        gd = (fim.fb_graddiv_env if self.env else fim.fb_graddiv)(vec_fb, *a["FBDiff"])
        gd = fim.omp_mult_vec(gd, a["PoissFact"])
        ee = fim.omp_add_vec(ee, gd)
        gg = dt_op * ee
        e_new = np.cos(w * x_focus) * ee + np.sin(w * x_focus) / w * gg
        gg = -w * np.sin(w * x_focus) * ee + np.cos(w * x_focus) * gg
        shift = np.exp(1j * kx_g[:, :, None, None] * x_focus)
        out = self.zeros_fb(6) if EG_fb is None else EG_fb
        out[:, :, :, :3] += e_new * shift
        out[:, :, :, 3:] += gg * shift
        return out

    # -- static-solution tables (solvers.py:348-358) -------------------------------------------
    def static_coeffs(self, px0):
        a = self.Args
        w = a["w"]
        beta0 = px0 / np.sqrt(1 + px0 ** 2)
        kxb = beta0 * a["kx_g"][:, :, None]
        c1 = np.zeros(w.shape + (2,), dtype="complex", order="F")
        c2 = np.zeros_like(c1)
        den = w ** 2 - kxb ** 2
        c1[..., 0], c1[..., 1] = 1.0j * kxb / den, -1.0 / den
        c2[..., 0], c2[..., 1] = w ** 2 / den, 1.0j * kxb / den
        return c1, c2

    # -- band-pass / anti-echo filter of the envelope solver (solvers.py:660-715) --------------
    def _antialias(self):
        a = self.Args
        kx_env, kx, kx0, nx = a["kx_env"], a["kx"], a["kx0"], a["Nx"]
        feats = a["Features"]
        cut = 0.85
        x = np.abs(kx_env) / np.abs(kx_env.max())
        band = ((x < cut) + (x >= cut) * np.cos(np.pi / 2 * (x - cut) / (1 - cut)) ** 2)[:, None, None]
        anti = np.ones_like(band)
        if "NoAntiEcho" not in feats:
            strength = feats["AntiEchoStrength"] if "AntiEchoStrength" in feats else 2
            n_echo = int(np.abs(kx).max() / np.abs(kx_env).max()) + 1
            echoes = np.abs(kx_env).max() / kx0 * np.arange(n_echo) - 1.0
            lo, hi = kx.min() / kx0 - 1.0, kx.max() / kx0 - 1.0
            hit = 0
            for pos in echoes:
                if not (lo < pos < hi):
                    continue
                s_loc = strength[hit] if isinstance(strength, (list, tuple)) else strength
                hit += 1
                if s_loc <= 0:
                    continue
                with np.errstate(divide="ignore", invalid="ignore"):  # the echo at pos = -1 has zero width, as in the reference
                    anti *= (1 - np.exp(-((kx / kx0 - 1 - pos) ** 2) / (s_loc * (pos + 1.0) / nx) ** 2))[:, None, None]
        return anti * band

    # -- convenience ---------------------------------------------------------------------------
    @staticmethod
    def get_damp_profile(lf):
        """``Solver.get_damp_profile`` (solvers.py:612-617): cos^2 ramp over the last quarter of ``int(lf)`` cells."""
        n = int(lf)
        g = np.arange(n)
        return (g >= 0.75 * n) * (0.5 - 0.5 * np.cos(np.pi * (g - 0.75 * n) / (0.25 * n))) ** 2

    def zeros_sp(self, ncomp=None):
        shp = self.shape_sp + ((ncomp,) if ncomp else ())
        return np.zeros(shp, dtype=complex, order="F")

    def zeros_fb(self, ncomp=None):
        shp = self.shape_fb + ((ncomp,) if ncomp else ())
        return np.zeros(shp, dtype=complex, order="F")
