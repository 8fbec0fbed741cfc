"""Second, independent restatement of the real-solver hot path in vectorised numpy (TEST INFRASTRUCTURE).

oracle/chimera_oracle.cpp restates the reference's Fortran loop for loop.  Nothing the reference ships
pins that restatement (no golden vectors, Fortran not buildable here), so this module restates the same
mathematics a second time in a *different form* -- whole-array numpy: ``np.add.at`` scatter for the
deposition, fancy-index gather, ``einsum``/``matmul`` + ``numpy.fft`` for the transforms, matrix notation
for the spectral calculus -- written from the Fortran sources cited per function.  tests/test_np_ref.py
requires the two restatements to agree to round-off; a transcription slip in either shows up there.

Both solver families are covered: the real ("PIC") one and, at the end of the file, the envelope ("KxShift") one
(grid_deps_env.f90, fb_math_env.f90), plus the devices, the SR integrals and the utils.f90 helpers.  Array
conventions as in the reference: Fortran order, grids (Nx, Nr, M[, c]) with radial node 0 the r = -dr/2 ghost.
"""
import numpy as np


# ---- f90/particle_tools.f90 -------------------------------------------------------------------
def push_velocs(momenta, fld, dt):
    """Boris rotation, particle_tools.f90:18-56 (dt already carries 2 pi q/m)."""
    p = np.array(momenta, dtype=float)
    e, b = 0.5 * dt * fld[:3], fld[3:]
    um = p + e
    g = np.sqrt(1.0 + (um ** 2).sum(0))
    t = 0.5 * dt * b / g
    s = 2.0 * t / (1.0 + (t ** 2).sum(0))
    u0 = um + np.cross(um, t, axis=0)
    up = um + np.cross(u0, s, axis=0)
    return up + e


def push_coords(coord, momenta, dt):
    """leap-frog, particle_tools.f90:58-82: returns (new coord, centred coord)."""
    g = np.sqrt(1.0 + (momenta ** 2).sum(0))
    new = coord + dt * momenta / g
    return new, 0.5 * (coord + new)


# ---- f90/grid_deps.f90 ------------------------------------------------------------------------
def _shape(coord, wghts, leftX, Rgrid, dx_inv, dr_inv):
    x, y, z = coord
    r = np.sqrt(y * y + z * z)
    ok = (wghts != 0.0) & (r < Rgrid[-1])
    ix = np.floor((x - leftX) * dx_inv).astype(int)
    ir = np.floor((r - Rgrid[0]) * dr_inv).astype(int)
    sx = (x - leftX) * dx_inv - ix
    sr = (r - Rgrid[np.clip(ir, 0, len(Rgrid) - 1)]) * dr_inv
    with np.errstate(invalid="ignore", divide="ignore"):
        ph = np.where(r > 0, (y - 1j * z) / np.where(r > 0, r, 1.0), 0.0)  # exp(-i theta); 0 on the axis
    return ok, ix, ir, sx, sr, ph


def _scatter(grid, ok, ix, ir, sx, sr, ph, amp, keep=None):
    """grid(Nx,Nr,M) += amp * exp(-i m theta) * Sx * Sr on the 4 nodes of each particle's cell.
    keep(i): optional per-particle mask of the x node ix + i (the chunk-edge rule of the *_chnk variants)."""
    nm = grid.shape[2]
    for m in range(nm):
        v = amp * ph ** m if m else amp.astype(complex)
        for i, wx in ((0, 1.0 - sx), (1, sx)):
            sel = ok if keep is None else ok & keep(i)
            for k, wr in ((0, 1.0 - sr), (1, sr)):
                np.add.at(grid[:, :, m], (ix[sel] + i, ir[sel] + k), (v * wx * wr)[sel])


def _fold_ghost(grid):
    grid[:, 1] -= grid[:, 0]
    grid[:, 0] = 0.0


def dep_dens(coord, wghts, dens, leftX, Rgrid, dx_inv, dr_inv, keep=None):
    """grid_deps.f90:89-147"""
    ok, ix, ir, sx, sr, ph = _shape(coord, wghts, leftX, Rgrid, dx_inv, dr_inv)
    _scatter(dens, ok, ix, ir, sx, sr, ph, wghts, keep and keep(ix))
    _fold_ghost(dens)
    return dens


def dep_curr(coord, momenta, wghts, curr, leftX, Rgrid, dx_inv, dr_inv, keep=None):
    """grid_deps.f90:18-87"""
    ok, ix, ir, sx, sr, ph = _shape(coord, wghts, leftX, Rgrid, dx_inv, dr_inv)
    ok = ok & (np.abs(momenta).sum(0) != 0.0)
    g = np.sqrt(1.0 + (momenta ** 2).sum(0))
    for l in range(3):
        _scatter(curr[..., l], ok, ix, ir, sx, sr, ph, momenta[l] * wghts / g, keep and keep(ix))
        _fold_ghost(curr[..., l])
    return curr


def chunk_rule(ind_in_chunk, guards, nxn):
    """The x-chunked variants (grid_deps_chnk.f90:38-124, grid_deps_env_chnk.f90) as a node mask on top of the plain
    deposit: particle ip belongs to chunk c (IndInChunk(c) <= ip < IndInChunk(c+1)) which owns the nodes
    [c cs, (c+1) cs), cs = Nx/nchnk.  A contribution to local node l = gx - c cs
      * with 0 < l < cs goes straight to the grid;
      * with l <= 0 goes to the chunk's `loc_left`, added back only if c cs - guards >= 0 (Q3: lost for chunk 0);
      * with l >= cs goes to `loc_right`, added back only if (c+1) cs + guards <= Nx - 1.
    Returns keep(ix) -> (i -> mask of the x node ix + i)."""
    nchnk = len(ind_in_chunk) - 1
    cs = nxn // nchnk

    def for_cells(ix):
        ip = np.arange(ix.size)
        c = np.searchsorted(np.asarray(ind_in_chunk)[1:], ip, side="right")
        inside = ip < ind_in_chunk[-1]
        c = np.minimum(c, nchnk - 1)
        left = c * cs

        def node(i):
            l = ix + i - left
            direct = (l > 0) & (l < cs)
            to_left = (l <= 0) & (l >= -guards) & (left - guards >= 0)
            to_right = (l >= cs) & (l <= cs + guards) & (left + cs + guards <= nxn - 1)
            return inside & (direct | to_left | to_right)

        return node

    return for_cells


def dep_dens_chnk(coord, wghts, dens, ind, guards, leftX, Rgrid, dx_inv, dr_inv):
    return dep_dens(coord, wghts, dens, leftX, Rgrid, dx_inv, dr_inv, keep=chunk_rule(ind, guards, dens.shape[0]))


def dep_curr_chnk(coord, momenta, wghts, curr, ind, guards, leftX, Rgrid, dx_inv, dr_inv):
    return dep_curr(coord, momenta, wghts, curr, leftX, Rgrid, dx_inv, dr_inv, keep=chunk_rule(ind, guards, curr.shape[0]))


def proj_fld(coord, wghts, Fld, Fld_tot, leftX, Rgrid, dx_inv, dr_inv):
    """grid_deps.f90:149-217 (gather phase exp(+i theta), 0 on the axis)"""
    ok, ix, ir, sx, sr, ph = _shape(coord, wghts, leftX, Rgrid, dx_inv, dr_inv)
    ph = np.conj(ph)
    out = np.array(Fld_tot, dtype=float)
    nm = Fld.shape[2]
    idx = np.nonzero(ok)[0]
    for l in range(6):
        acc = np.zeros(idx.size)
        for m in range(nm):
            pm = ph[idx] ** m if m else np.ones(idx.size, dtype=complex)
            for i, wx in ((0, 1.0 - sx[idx]), (1, sx[idx])):
                for k, wr in ((0, 1.0 - sr[idx]), (1, sr[idx])):
                    acc += (wx * wr * pm * Fld[ix[idx] + i, ir[idx] + k, m, l]).real
        out[l, idx] += acc
    return out


def eb_correction(eb):
    """grid_deps.f90:219-266"""
    eb = np.array(eb)
    eb[:, :, 0] /= 2 * np.pi
    eb[:, :, 1:] /= np.pi
    eb[:, 0, 0] = eb[:, 1, 0]
    eb[:, 0, 1:] = -eb[:, 1, 1:]
    return eb


# ---- f90/fb_io.f90 ----------------------------------------------------------------------------
def fb_in(vec, leftX, kx, In):
    """fb_vec_in / fb_scl_in, fb_io.f90:18-98: DHT over r (ghost node skipped), FFT over x, phase."""
    a = np.einsum("xrm...,rkm->xkm...", vec[:, 1:], In)
    a = np.fft.fft(a, axis=0)
    ph = np.exp(-1j * leftX * kx)
    return a * ph.reshape((-1,) + (1,) * (a.ndim - 1))


def fb_out(vec_fb, leftX, kx, Out):
    """fb_vec_out / fb_scl_out, fb_io.f90:100-180: DHT, phase, unnormalised inverse FFT; ghost row 0."""
    a = np.einsum("xkm...,krm->xrm...", vec_fb, Out)
    ph = np.exp(1j * leftX * kx)
    a = np.fft.ifft(a * ph.reshape((-1,) + (1,) * (a.ndim - 1)), axis=0) * vec_fb.shape[0]
    out = np.zeros((a.shape[0], a.shape[1] + 1) + a.shape[2:], dtype=complex)
    out[:, 1:] = a
    return out


def fb_eb_out(e_fb, b_fb, leftX, kx, Out):
    """fb_io.f90:182-228"""
    return np.concatenate((fb_out(e_fb[..., :3], leftX, kx, Out), fb_out(b_fb, leftX, kx, Out)), axis=-1)


# ---- f90/fb_math.f90 --------------------------------------------------------------------------
MIRROR_SHIFT = 0  # 0: rows are a full kx axis (or the rank-0 slab); 1: a rank>0 kx slab (chimera_b200/sharding.py)


def _mirror(f):
    """-conj(f(-kx)): rows (Nx - i) mod Nx, fb_math.f90:35-36 (slab-local form: (L - i - shift) mod L)"""
    n = f.shape[0]
    return -np.conj(f[(n - np.arange(n) - MIRROR_SHIFT) % n])


def _con(D, f):
    """(D . f)[x, k'] = sum_k D[k, k'] f[x, k]"""
    return f @ D


def _lower(f, m):
    """the mode-(m-1) neighbour of slot m; for m = 0 the mirrored slot 1 (0 when there is none: Q7)"""
    if m > 0:
        return f[:, :, m - 1]
    return _mirror(f[:, :, 1]) if f.shape[2] > 1 else np.zeros_like(f[:, :, 0])


def fb_grad(scl, Dp, Dm, kx):
    """fb_math.f90:96-149"""
    nm = scl.shape[2]
    out = np.zeros(scl.shape + (3,), dtype=complex)
    ikx = 1j * kx[:, None]
    for m in range(nm):
        out[:, :, m, 0] = ikx * scl[:, :, m]
        t = _con(Dm[:, :, m], _lower(scl, m))
        out[:, :, m, 1] -= t
        out[:, :, m, 2] += 1j * t
        if m < nm - 1:
            t = _con(Dp[:, :, m], scl[:, :, m + 1])
            out[:, :, m, 1] += t
            out[:, :, m, 2] += 1j * t
    return out


def fb_div(vec, Dp, Dm, kx, extra_mode=False):
    """fb_math.f90:151-199; extra_mode: the (M+1)-slot scalar fb_graddiv builds internally (:232-262)"""
    nm = vec.shape[2]
    n_out = nm + 1 if extra_mode else nm
    out = np.zeros(vec.shape[:2] + (n_out,), dtype=complex)
    ikx = 1j * kx[:, None]
    y, z = vec[..., 1], vec[..., 2]
    for m in range(n_out):
        if m < nm:
            out[:, :, m] += ikx * vec[:, :, m, 0]
        out[:, :, m] += _con(Dm[:, :, m], 1j * _lower(z, m) - _lower(y, m))
        if m < nm - 1:
            out[:, :, m] += _con(Dp[:, :, m], 1j * z[:, :, m + 1] + y[:, :, m + 1])
    return out


def fb_rot(vec, Dp, Dm, kx):
    """fb_math.f90:18-94"""
    nm = vec.shape[2]
    out = np.zeros_like(vec)
    ikx = 1j * kx[:, None]
    x, y, z = vec[..., 0], vec[..., 1], vec[..., 2]
    for m in range(nm):
        out[:, :, m, 1] -= ikx * z[:, :, m]
        out[:, :, m, 2] += ikx * y[:, :, m]
        if m < nm - 1:
            out[:, :, m, 0] -= _con(Dp[:, :, m], 1j * y[:, :, m + 1] - z[:, :, m + 1])
            t = _con(Dp[:, :, m], x[:, :, m + 1])
            out[:, :, m, 1] += 1j * t
            out[:, :, m, 2] -= t
        out[:, :, m, 0] -= _con(Dm[:, :, m], 1j * _lower(y, m) + _lower(z, m))
        t = _con(Dm[:, :, m], _lower(x, m))
        out[:, :, m, 1] += 1j * t
        out[:, :, m, 2] += t
    return out


def fb_graddiv(vec, Dp, Dm, kx):
    """fb_math.f90:201-293: grad of the (M+1)-slot divergence; the mirror terms only when M > 1 (:217)"""
    nm = vec.shape[2]
    s = fb_div(vec, Dp, Dm, kx, extra_mode=True)
    out = np.zeros_like(vec)
    ikx = 1j * kx[:, None]
    for m in range(nm):
        out[:, :, m, 0] = ikx * s[:, :, m]
        low = s[:, :, m - 1] if m > 0 else (_mirror(s[:, :, 1]) if nm > 1 else np.zeros_like(s[:, :, 0]))
        t = _con(Dm[:, :, m], low)
        out[:, :, m, 1] -= t
        out[:, :, m, 2] += 1j * t
        t = _con(Dp[:, :, m], s[:, :, m + 1])
        out[:, :, m, 1] += t
        out[:, :, m, 2] += 1j * t
    return out


# ---- f90/maxwell_solvers.f90 --------------------------------------------------------------------
def maxwell_push_with_spchrg(EG, J, g_n, g_np1, C1, C2):
    """maxwell_solvers.f90:18-60"""
    E, G = EG[..., :3], EG[..., 3:]
    src = (E, G, J, g_n, g_np1)
    newE = sum(C1[..., i, None] * s for i, s in enumerate(src))
    newG = sum(C2[..., i, None] * s for i, s in enumerate(src))
    return np.concatenate((newE, newG), axis=-1)


def maxwell_push_wo_spchrg(EG, J, C1, C2):
    """maxwell_solvers.f90:62-96"""
    E, G = EG[..., :3], EG[..., 3:]
    src = (E, G, J)
    newE = sum(C1[..., i, None] * s for i, s in enumerate(src))
    newG = sum(C2[..., i, None] * s for i, s in enumerate(src))
    return np.concatenate((newE, newG), axis=-1)


def poiss_corr(J, gdj, g_n, g_np1, dt_inv, w2_inv):
    """maxwell_solvers.f90:131-164"""
    return J + w2_inv[..., None] * (gdj + dt_inv * (g_np1 - g_n))


# ------------------------------------------------------------------ devices.f90 (NEXT-1), whole-array form
def _ramp(x, X0, Lx, ramp):
    a = np.ones_like(x)
    a = np.where((x > X0 + Lx - ramp) & (x < X0 + Lx), (X0 + Lx - x) / ramp, a)
    a = np.where((x > X0) & (x < X0 + ramp), (x - X0) / ramp, a)
    return np.where((x <= X0) | (x >= X0 + Lx), 0.0, a)


def undul_analytic(coord, fld, t, params, taper=None):
    a0, lam, X0, Lx = params[:4]
    ku = 2 * np.pi / lam
    x, y = coord[0], coord[1]
    ampl = _ramp(x, X0, Lx, lam)
    if taper is not None:
        ampl = ampl * (1 + taper * (x - X0 - 0.5 * Lx) / (0.5 * Lx))
    ampl = ampl * a0
    out = fld.copy()
    out[4] += ampl * np.sin(ku * (x - X0)) * np.cosh(ku * y)
    out[3] += ampl * np.cos(ku * (x - X0)) * np.sinh(ku * y)
    return out


def undul_analytic_taper(coord, fld, t, params):
    return undul_analytic(coord, fld, t, params, taper=params[4])


def undul_mapped(coord, fld, t, a0, params, tap=False):
    lam, Xleft, dx = params[:3]
    nx = a0.shape[1]
    ku = 2 * np.pi / lam
    x, y = coord[0], coord[1]
    ok = ~((x < Xleft + dx) | (x > Xleft + nx * dx - dx))
    s = (x - Xleft) / dx
    ix = np.floor(s + 0.5).astype(np.int64)
    d = s - ix
    S0 = np.stack((0.5 * (0.5 - d) ** 2, 0.75 - d ** 2, 0.5 * (0.5 + d) ** 2))
    pad = np.zeros((2, nx + 2))  # column k = node k (1-based); node 0 and nx+1 zero (Q12)
    pad[:, 1:nx + 1] = a0
    k = np.clip(ix[None, :] + np.arange(-1, 2)[:, None], 0, nx + 1)
    s1 = (S0 * pad[0][k]).sum(0)
    s2 = (S0 * pad[1][k]).sum(0)
    amp = 1.0
    if tap:
        Lx, taper = params[3], params[4]
        amp = 1 + taper * (x - 0.5 * (nx * dx - Lx) - Xleft - 0.5 * Lx) / (0.5 * Lx)
    out = fld.copy()
    out[4] += np.where(ok, amp * s1 * np.cosh(ku * y), 0.0)
    out[3] += np.where(ok, amp * s2 * np.sinh(ku * y), 0.0)
    return out


def undul_mapped_tap(coord, fld, t, a0, params):
    return undul_mapped(coord, fld, t, a0, params, tap=True)


def planewave(coord, fld, t, params):
    a0, lam, X0, Lx, ramp, theta, phi0 = params
    x, y = coord[0], coord[1]
    k0 = 2 * np.pi / lam
    ampl = a0 * _ramp(x, X0, Lx, ramp) * np.sin(k0 * (x * np.cos(theta) + y * np.sin(theta) - t) + phi0)
    out = fld.copy()
    out[2] += ampl
    out[3] += ampl * np.sin(theta)
    out[4] -= ampl * np.cos(theta)
    return out


def gaussbeam(coord, fld, t, a0, params):
    lam, axis, x0, y0, z0, Lx, Ly, Lz = params
    xp, yp, zp = coord[0] - x0 - axis * t, coord[1] - y0, coord[2] - z0
    E = a0 * np.exp(-xp * xp / Lx ** 2 - yp * yp / Ly ** 2 - zp * zp / Lz ** 2) * np.sin(2 * np.pi / lam * xp)
    out = fld.copy()
    out[2] += E
    out[4] -= axis * E
    return out


# ---- f90/SR.f90 --------------------------------------------------------------------------------
# Whole-array form: all (time, particle, pixel) phases and amplitudes at once, the frequency axis last;
# tracks are (3, nt, np) as in moduls/SR.py:141-151.
def _guarded_integral(amp, phase, limit, omega):
    """sum over time of amp * exp(i omega phase) where omega*|phase(it) - phase(it-1)| < limit
    (phase(0) := 0: the reference starts C3_prev / phase_prv at zero, SR.f90:65,302).
    amp (..., nt) complex-able, phase (..., nt) -> (..., nom)"""
    prev = np.concatenate([np.zeros_like(phase[..., :1]), phase[..., :-1]], axis=-1)
    dph = np.abs(phase - prev)
    keep = dph[..., None] * omega < limit                      # (..., nt, nom)
    e = np.exp(1j * phase[..., None] * omega)
    return (np.where(keep, e, 0.0) * amp[..., None]).sum(-2)


def sr_calc_far(spect, coords, mom_prv, mom_nxt, wghts, dt, omega, sin_th, cos_th, sin_ph, cos_ph, comp=0):
    """SR.f90:18-137 (comp=0) / :139-254 (comp=1..3): far-field Lienard-Wiechert spectrum
    d2W/dOmega domega ~ |int n x ((n - beta) x beta') / (1 - n.beta)^2 e^{i omega (t - n.r)} dt|^2,
    in the reference's component form C4 = (C1 (n - beta) - C2 beta') / C2^2."""
    x = np.moveaxis(np.asarray(coords), 0, -1)[:, :, ::-1].transpose(1, 0, 2)   # (np, nt, [z,y,x] -> reversed)
    bp = np.asarray(mom_prv) / np.sqrt(1.0 + (np.asarray(mom_prv) ** 2).sum(0))
    bn = np.asarray(mom_nxt) / np.sqrt(1.0 + (np.asarray(mom_nxt) ** 2).sum(0))
    acc = np.moveaxis((bn - bp) / dt, 0, -1).transpose(1, 0, 2)                 # (np, nt, 3) index 0 = longitudinal
    vel = np.moveaxis(0.5 * (bn + bp), 0, -1).transpose(1, 0, 2)
    xx = np.moveaxis(np.asarray(coords), 0, -1).transpose(1, 0, 2)
    nt = xx.shape[1]
    tt = np.arange(1, nt + 1) * dt
    out = np.array(spect, dtype=float, order="F")
    for iph in range(len(sin_ph)):
        for ith in range(len(sin_th)):
            n = np.array([cos_th[ith], sin_th[ith] * sin_ph[iph], sin_th[ith] * cos_ph[iph]])  # (x, y, z) of the reference
            c2 = 1.0 - vel @ n
            c1 = acc @ n
            c3 = 2.0 * np.pi * (tt[None, :] - xx @ n)
            c4 = (c1[..., None] * (n - vel) - c2[..., None] * acc) / c2[..., None] ** 2 * dt   # (np, nt, 3)
            comps = range(3) if comp == 0 else ([comp - 1] if 1 <= comp <= 3 else [])
            tot = np.zeros((xx.shape[0], len(omega)))
            for k in comps:
                tot += np.abs(_guarded_integral(c4[..., k], c3, np.pi, np.asarray(omega))) ** 2
            out[:, ith, iph] += (np.abs(wghts)[:, None] * tot).sum(0)
    return out


def sr_calc_near(spect, coords, mom, wghts, dt, omega, g1, g2, z_scr, comp=0, circ=None):
    """SR.f90:256-447 (Cartesian screen: g1 = X, g2 = Y) and :449-642 (polar: g1 = R, circ = (SinPh, CosPh)).
    Integrand (i omega (beta - n)/R + n / (2 pi R^2)) dt e^{2 pi i omega (t + R)}; the screen x pairs with the
    third track coordinate, y with the second, z_scr with the first (:305-307)."""
    xx = np.moveaxis(np.asarray(coords), 0, -1).transpose(1, 0, 2)   # (np, nt, 3)
    u = np.asarray(mom)
    vel = np.moveaxis(u / np.sqrt(1.0 + (u ** 2).sum(0)), 0, -1).transpose(1, 0, 2)
    nt = xx.shape[1]
    tt = np.arange(1, nt + 1) * dt
    out = np.array(spect, dtype=float, order="F")
    n1, n2 = out.shape[1:]
    om = np.asarray(omega)
    for i1 in range(n1):
        for i2 in range(n2):
            if circ is None:
                xs, ys = g1[i1], g2[i2]
            else:
                xs, ys = g1[i1] * circ[1][i2], g1[i1] * circ[0][i2]
            d = np.array([z_scr, ys, xs]) - xx
            r0 = np.sqrt((d ** 2).sum(-1))
            n = d / r0[..., None]
            phase = 2.0 * np.pi * (tt[None, :] + r0)
            a1 = dt / r0[..., None] * (vel - n)
            a2 = dt / r0[..., None] ** 2 / (2.0 * np.pi) * n
            prev = np.concatenate([np.zeros_like(phase[:, :1]), phase[:, :-1]], axis=1)
            keep = np.abs(phase - prev)[..., None] * om < 2.0 * np.pi
            e = np.where(keep, np.exp(1j * phase[..., None] * om), 0.0)          # (np, nt, nom)
            comps = range(3) if comp == 0 else [comp - 1]
            tot = np.zeros((xx.shape[0], len(om)))
            for k in comps:
                integ = ((1j * a1[..., k, None] * om + a2[..., k, None]) * e).sum(1)
                tot += np.abs(integ) ** 2
            out[:, i1, i2] += (np.abs(wghts)[:, None] * tot).sum(0)
    return out


# ---- f90/utils.f90 -----------------------------------------------------------------------------
def intens_profo(fld, no):
    """utils.f90:18-57: |sum_m e^{i m theta} F_m|^2 summed over x and the 3 components, on `no` angles
    theta_j = 2 pi j/(no-1); mode slots ordered -nko..nko; ghost radial node dropped."""
    nm = fld.shape[2]
    nko = (nm - 1) // 2
    theta = 2.0 * np.pi * np.arange(no) / (no - 1.0)
    ph = np.exp(1j * np.arange(-nko, nko + 1)[None, :] * theta[:, None])      # (no, nm)
    s = np.einsum("om,xrml->oxrl", ph, fld[:, 1:])
    return (np.abs(s) ** 2).sum((1, 3))


def density_2x(x, y, wght, grid, bins_x, bins_y):
    """utils.f90:210-272: weighted 2-D histogram with the reference's 5-node shape
    S = (|d^3|/3 [d<0], 1/4 - d/2 + d^3/3, 1/2 - |d^3|/3, 1/4 + d/2 - d^3/3, |d^3|/3 [d>=0])."""
    dxg, dyg = (grid[1] - grid[0]) / bins_x, (grid[3] - grid[2]) / bins_y
    sel = (x >= grid[0]) & (x <= grid[0] + dxg * bins_x) & (y >= grid[2]) & (y <= grid[2] + dyg * bins_y)
    x, y, w = x[sel], y[sel], wght[sel]

    def shape(u, orig, dlt):
        k = np.floor((u - orig) / dlt + 0.5).astype(int)
        d = (u - (dlt * k + orig)) / dlt
        d3 = d ** 3
        s = np.stack([np.where(d < 0, np.abs(d3) / 3, 0.0), 0.25 - 0.5 * d + d3 / 3, 0.5 - np.abs(d3) / 3,
                      0.25 + 0.5 * d - d3 / 3, np.where(d >= 0, np.abs(d3) / 3, 0.0)])
        return k, s

    kx, sx = shape(x, grid[0], dxg)
    ky, sy = shape(y, grid[2], dyg)
    dens = np.zeros((bins_x + 5, bins_y + 5), order="F")
    for i in range(5):
        for j in range(5):
            np.add.at(dens, (kx + i, ky + j), w * sx[i] * sy[j])
    return dens / dxg / dyg


# ---- envelope ("KxShift") family: f90/grid_deps_env.f90, f90/fb_math_env.f90 -------------------------
# Mode slots are ordered -nko..nko (slot j <-> mode j - nko); the operators DpS2S / DmS2S carry 2 nko + 3 slots,
# -nko-1..nko+1 (slot j <-> mode j - nko - 1).  Everything is written per MODE here, not per slot.
def _env_modes(nm):
    nko = (nm - 1) // 2
    return nko, range(-nko, nko + 1)


def _carrier(x, kx0, sign):
    return np.cos(x * kx0) + sign * 1j * np.sin(x * kx0)


def _scatter_env(grid, ok, ix, ir, sx, sr, ph, amp, keep=None):
    """grid(Nx,Nr,2nko+1) += amp * exp(-i m theta) * Sx * Sr for m = -nko..nko (exp(-i|m|theta) conjugated for m<0)"""
    nko, modes = _env_modes(grid.shape[2])
    for m in modes:
        pm = ph ** abs(m) if m else np.ones_like(ph)
        v = amp * (np.conj(pm) if m < 0 else pm)
        for i, wx in ((0, 1.0 - sx), (1, sx)):
            sel = ok if keep is None else ok & keep(i)
            for k, wr in ((0, 1.0 - sr), (1, sr)):
                np.add.at(grid[:, :, m + nko], (ix[sel] + i, ir[sel] + k), (v * wx * wr)[sel])


def dep_dens_env(coord, wghts, dens, leftX, Rgrid, dx_inv, dr_inv, kx0, keep=None):
    """grid_deps_env.f90:96-162.  Q2: the complex weight w e^{-i kx0 x} enters SQUARED (:145,147)"""
    ok, ix, ir, sx, sr, ph = _shape(coord, wghts, leftX, Rgrid, dx_inv, dr_inv)
    wc = wghts * _carrier(coord[0], kx0, -1)
    _scatter_env(dens, ok, ix, ir, sx, sr, ph, wc * wc, keep and keep(ix))
    _fold_ghost(dens)
    return dens


def dep_curr_env(coord, momenta, wghts, curr, leftX, Rgrid, dx_inv, dr_inv, kx0, keep=None):
    """grid_deps_env.f90:18-94.  Q1: only the third component is deposited (:76)"""
    ok, ix, ir, sx, sr, ph = _shape(coord, wghts, leftX, Rgrid, dx_inv, dr_inv)
    ok = ok & (np.abs(momenta).sum(0) != 0.0)
    g = np.sqrt(1.0 + (momenta ** 2).sum(0))
    _scatter_env(curr[..., 2], ok, ix, ir, sx, sr, ph, wghts * _carrier(coord[0], kx0, -1) * momenta[2] / g,
                 keep and keep(ix))
    for l in range(3):
        _fold_ghost(curr[..., l])
    return curr


def proj_fld_env(coord, wghts, Fld, Fld_tot, leftX, Rgrid, dx_inv, dr_inv, kx0):
    """grid_deps_env.f90:164-238: Re[ e^{+i kx0 x} e^{+i m theta} Sx Sr F_m ], phase 1 on the axis (Q4)"""
    ok, ix, ir, sx, sr, ph = _shape(coord, wghts, leftX, Rgrid, dx_inv, dr_inv)
    r = np.sqrt(coord[1] ** 2 + coord[2] ** 2)
    ph = np.where(r > 0, np.conj(ph), 1.0)
    car = _carrier(coord[0], kx0, +1)
    out = np.array(Fld_tot, dtype=float)
    nko, modes = _env_modes(Fld.shape[2])
    idx = np.nonzero(ok)[0]
    for l in range(6):
        acc = np.zeros(idx.size)
        for m in modes:
            pm = ph[idx] ** abs(m) if m else np.ones(idx.size, dtype=complex)
            pm = car[idx] * (np.conj(pm) if m < 0 else pm)
            for i, wx in ((0, 1.0 - sx[idx]), (1, sx[idx])):
                for k, wr in ((0, 1.0 - sr[idx]), (1, sr[idx])):
                    acc += (wx * wr * pm * Fld[ix[idx] + i, ir[idx] + k, m + nko, l]).real
        out[l, idx] += acc
    return out


def eb_correction_env(eb):
    """grid_deps_env.f90:240-283: 1/pi for every mode; ghost row copied (nko = 0) or negated for ALL modes (Q6)"""
    eb = np.array(eb) / np.pi
    eb[:, 0] = eb[:, 1] if eb.shape[2] == 1 else -eb[:, 1]
    return eb


class _EnvOps:
    """per-mode access: f[m] = field of mode m (zero outside -nko..nko), Dp[m] / Dm[m] for m in -nko-1..nko+1"""

    def __init__(self, Dp, Dm, nm):
        self.nko = (nm - 1) // 2
        self.Dp, self.Dm = Dp, Dm

    def dp(self, m):
        return self.Dp[:, :, m + self.nko + 1]

    def dm(self, m):
        return self.Dm[:, :, m + self.nko + 1]


def fb_grad_env(scl, Dp, Dm, kx):
    """fb_math_env.f90:18-61"""
    o = _EnvOps(Dp, Dm, scl.shape[2])
    nko = o.nko
    out = np.zeros(scl.shape[:2] + (scl.shape[2], 3), dtype=complex)
    for m in range(-nko, nko + 1):
        j = m + nko
        out[:, :, j, 0] = 1j * kx[:, None] * scl[:, :, j]
        if m > -nko:
            g = scl[:, :, j - 1] @ o.dm(m)
            out[:, :, j, 1] -= g
            out[:, :, j, 2] += 1j * g
        if m < nko:
            g = scl[:, :, j + 1] @ o.dp(m)
            out[:, :, j, 1] += g
            out[:, :, j, 2] += 1j * g
    return out


def _div_env(vec, o, kx, m_lo, m_hi):
    """S[m] for m = m_lo..m_hi (fb_math_env.f90:63-104; :183-206 with the two extra modes)"""
    nko = o.nko
    S = np.zeros(vec.shape[:2] + (m_hi - m_lo + 1,), dtype=complex)
    for m in range(m_lo, m_hi + 1):
        s = np.zeros(vec.shape[:2], dtype=complex)
        if -nko <= m <= nko:
            s += 1j * kx[:, None] * vec[:, :, m + nko, 0]
        if m > -nko:
            s -= (vec[:, :, m - 1 + nko, 1] - 1j * vec[:, :, m - 1 + nko, 2]) @ o.dm(m)
        if m < nko:
            s += (vec[:, :, m + 1 + nko, 1] + 1j * vec[:, :, m + 1 + nko, 2]) @ o.dp(m)
        S[:, :, m - m_lo] = s
    return S


def fb_div_env(vec, Dp, Dm, kx):
    o = _EnvOps(Dp, Dm, vec.shape[2])
    return _div_env(vec, o, kx, -o.nko, o.nko)


def fb_rot_env(vec, Dp, Dm, kx):
    """fb_math_env.f90:106-162.  Q5: the Dm term of the first component is computed and dropped (:146-151)"""
    o = _EnvOps(Dp, Dm, vec.shape[2])
    nko = o.nko
    out = np.zeros_like(vec)
    ikx = 1j * kx[:, None]
    for m in range(-nko, nko + 1):
        j = m + nko
        out[:, :, j, 1] = -ikx * vec[:, :, j, 2]
        out[:, :, j, 2] = ikx * vec[:, :, j, 1]
        if m < nko:
            out[:, :, j, 0] -= (1j * vec[:, :, j + 1, 1] - vec[:, :, j + 1, 2]) @ o.dp(m)
            g = vec[:, :, j + 1, 0] @ o.dp(m)
            out[:, :, j, 1] += 1j * g
            out[:, :, j, 2] -= g
        if m > -nko:
            g = vec[:, :, j - 1, 0] @ o.dm(m)
            out[:, :, j, 1] += 1j * g
            out[:, :, j, 2] += g
    return out


def fb_graddiv_env(vec, Dp, Dm, kx):
    """fb_math_env.f90:164-233: divergence on modes -nko-1..nko+1, then the gradient with both neighbours always"""
    o = _EnvOps(Dp, Dm, vec.shape[2])
    nko = o.nko
    S = _div_env(vec, o, kx, -nko - 1, nko + 1)  # S[:, :, m + nko + 1]
    out = np.zeros_like(vec)
    for m in range(-nko, nko + 1):
        j = m + nko
        out[:, :, j, 0] = 1j * kx[:, None] * S[:, :, m + nko + 1]
        g1 = S[:, :, m - 1 + nko + 1] @ o.dm(m)
        g2 = S[:, :, m + 1 + nko + 1] @ o.dp(m)
        out[:, :, j, 1] = -g1 + g2
        out[:, :, j, 2] = 1j * g1 + 1j * g2
    return out


def dep_dens_env_chnk(coord, wghts, dens, ind, guards, leftX, Rgrid, dx_inv, dr_inv, kx0):
    return dep_dens_env(coord, wghts, dens, leftX, Rgrid, dx_inv, dr_inv, kx0, keep=chunk_rule(ind, guards, dens.shape[0]))


def dep_curr_env_chnk(coord, momenta, wghts, curr, ind, guards, leftX, Rgrid, dx_inv, dr_inv, kx0):
    return dep_curr_env(coord, momenta, wghts, curr, leftX, Rgrid, dx_inv, dr_inv, kx0,
                        keep=chunk_rule(ind, guards, curr.shape[0]))


# ---- f90/particle_tools.f90: generation, culling, chunking, permutation --------------------------------
def genparts(Xgrid, Rgrid, RandPackO, PackX, PackR, PackO):
    """particle_tools.f90:84-128: PPC particles per (x, r) cell at x0 + (x1-x0) PackX, r0 + (r1-r0) PackR on the
    half-cell-shifted r grid, azimuth PackO * exp(2 pi i RandPackO(cell)); particles with r <= 0 are skipped.
    Returns coord(4, n) = (x, r sin, r cos, r) in the reference's order (r outer loop, x inner, then the pack)."""
    dr_2 = 0.5 * (Rgrid[1] - Rgrid[0])
    rs = Rgrid + dr_2
    x0, x1 = Xgrid[:-1], Xgrid[1:]
    r0, r1 = rs[:-1], rs[1:]
    xx = x0[None, :, None] + (x1 - x0)[None, :, None] * PackX[None, None, :]              # (1, nx-1, ppc)
    rr = r0[:, None, None] + (r1 - r0)[:, None, None] * PackR[None, None, :]              # (nr-1, 1, ppc)
    ang = 2.0 * np.pi * RandPackO[:len(x0), :len(r0)].T[:, :, None]
    oc = PackO[None, None, :] * (np.cos(ang) + 1j * np.sin(ang))                          # (nr-1, nx-1, ppc)
    xx, rr = np.broadcast_to(xx, oc.shape), np.broadcast_to(rr, oc.shape)
    keep = rr > 0
    return np.asfortranarray(np.stack((xx[keep], (rr * oc.imag)[keep], (rr * oc.real)[keep], rr[keep])))


def _inside(coord, lims):
    r2 = coord[1] ** 2 + coord[2] ** 2
    return (coord[0] >= lims[0]) & (coord[0] <= lims[1]) & (r2 >= lims[2]) & (r2 <= lims[3])


def sortpartsout(coord, lims):
    """particle_tools.f90:130-153: 0-based indices of the particles inside lims = (xmin, xmax, r2min, r2max)"""
    return np.nonzero(_inside(coord, lims))[0]


def sortoutghosts(coord):
    """particle_tools.f90:326-341: 0-based indices of the entries that are not exactly zero"""
    return np.nonzero(np.asarray(coord) != 0.0)[0]


def chunk_coords_boundaries(coord, lims, Xgrid, nchnk):
    """particle_tools.f90:155-208: chunk id per particle (-2 outside lims), prefix offsets, number leaving"""
    n = len(Xgrid)
    length = (Xgrid[n // nchnk] - Xgrid[0]) if nchnk > 1 else (Xgrid[-1] - Xgrid[0])
    ich = np.floor((coord[0] - Xgrid[0]) * (1.0 / length)).astype(int)
    ins = _inside(coord, lims)
    ids = np.where(ins, ich, -2).astype(np.int8)
    counts = np.bincount(ich[ins], minlength=nchnk)[:nchnk]
    return ids, np.concatenate(([0], np.cumsum(counts))).astype(np.int32), int((~ins).sum())


def align_data(dat, idx):
    """particle_tools.f90:270-324 (vec and scl): leading len(idx) entries <- dat[..., idx]"""
    out = np.array(dat)
    out[..., :len(idx)] = dat[..., idx]
    return out


# ---- f90/maxwell_solvers.f90: the rest of the elementwise family --------------------------------------
def maxwell_init_push(EG, J, g_n, C1, C2):
    """maxwell_solvers.f90:98-129"""
    out = np.array(EG)
    out[..., :3] += C1[..., 0:1] * J + C1[..., 1:2] * g_n
    out[..., 3:] += C2[..., 0:1] * J + C2[..., 1:2] * g_n
    return out


def poiss_corr_stat(J, gdj, g_n, DT, w2_inv):
    """maxwell_solvers.f90:166-197"""
    return J + (gdj + DT[:, None, None, None] * g_n) * w2_inv[..., None]


def field_drift(EG, kx, beta0, dt):
    """maxwell_solvers.f90:199-226"""
    return EG * np.exp(-0.5j * dt * beta0 * kx)[:, None, None, None]


# ---- f90/fb_io.f90:230-300 -----------------------------------------------------------------------------
def fb_filtr(vec, leftX, kx, filtr, modefilt=0):
    """x-space window on spectral fields: back to x (unnormalised inverse FFT after the phase), multiply the first
    len(filtr) nodes (mode 0, 'left', the only one the driver uses: solvers.py:619), forward FFT, undo phase and Nx"""
    nx = vec.shape[0]
    shift = (np.cos(leftX * kx) + 1j * np.sin(leftX * kx)).reshape((nx,) + (1,) * (vec.ndim - 1))
    a = np.fft.ifft(vec * shift, axis=0) * nx
    w = np.ones(nx)
    if modefilt in (0, 2):
        w[:len(filtr)] *= filtr
    if modefilt in (1, 2):  # Aifft(nkx-nxfilt:nkx) has nxfilt+1 elements against filtr's nxfilt: shape bug in the
        raise NotImplementedError("modes 1, 2 of fb_filtr are ill-formed in the reference (fb_io.f90:279) and unused")
    a *= w.reshape(shift.shape)
    return np.fft.fft(a, axis=0) / (nx * shift)
