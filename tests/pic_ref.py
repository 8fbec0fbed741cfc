"""Reference step sequence on top of any `fimera`-compatible module (TEST INFRASTRUCTURE).

A compact restatement of the reference's Python driver for the configurations the engine covers:
``ChimeraRun.make_halfstep`` / ``make_step`` (reference moduls/chimera_main.py:61-92), the solver
wrappers (moduls/solvers.py:281-331, 407-465, 517-553) and the species wrappers
(moduls/species.py:246-398).  It is run on the CPU oracle (``oracle.fimera``) to produce the expected
state for the device-resident engine, and on ``chimera_b200.fimera`` to exercise the host-buffer
drop-in path end to end.  tools/gen_golden.py checks this restatement against the reference's own,
unmodified driver (in the build container, where /root/reference exists).
"""
import numpy as np


class RefSpecies:
    def __init__(self, coords, momenta, weights, charge=-1.0, mass=1.0, still=False, device=None, devices=None):
        self.coords = np.asfortranarray(coords, dtype=float).copy(order="F")
        self.coords_halfstep = self.coords.copy(order="F")
        self.momenta = np.asfortranarray(momenta, dtype=float).copy(order="F")
        self.weights = np.asfortranarray(weights, dtype=float).copy()
        self.push_fact = 2 * np.pi * charge / mass  # species.py:64
        self.still = still
        # species.py:55 Args['Devices']: a list of (callable, *args); `device` is the one-entry short form
        self.devices = list(devices) if devices else ([tuple(device)] if device is not None else [])
        self.EB = np.zeros((6, 0), order="F")
        self.chunks = None


class RefRun:
    def __init__(self, fim, setup, species, chunked=None, sort_every=None, poisson_iters=None, background=False,
                 reduce=None, rank=0):
        self.f, self.S, self.sp = fim, setup, species
        a = self.a = setup.Args
        feats = a.get("Features", ())
        self.env = setup.env
        self.space_charge = "SpaceCharge" in feats
        self.static_kick = "StaticKick" in feats  # chimera_main.py:106-125, 186: quasi-static field of a beam
        self.chunked = ("Xchunked" in a) if chunked is None else chunked
        self.nchnk, self.guards = (a["Xchunked"] if self.chunked else (1, 0))
        self.sort_every = (self.guards + 1 if self.chunked else 0) if sort_every is None else sort_every
        self.npoiss = (0 if "NoPoissonCorrection" in feats else 3) if poisson_iters is None else poisson_iters
        self.background = background
        # multi-rank runs (particles sharded): `reduce` sums a deposited grid over the ranks in place
        self.reduce, self.rank = reduce, rank
        S = setup
        self.J, self.Rho, self.Bck = S.zeros_sp(3), S.zeros_sp(), S.zeros_sp()
        self.EB = S.zeros_sp(6)
        self.EG_fb, self.J_fb, self.B_fb = S.zeros_fb(6), S.zeros_fb(3), S.zeros_fb(3)
        self.Rho_fb, self.vec_fb = S.zeros_fb(), S.zeros_fb(3)
        self.g_prv, self.g_nxt = S.zeros_fb(3), S.zeros_fb(3)
        self.PE, self.PG = S.PSATD_E, S.PSATD_G
        self.istep = 0
        # a window that moves every step ('Steps': 1): (shift at stage 1, shift at stage 2), chimera_main.py:40-51
        self.window = (0.0, 0.0)

    # ---- species.py:351-398 -------------------------------------------------------------------
    def chunk_and_damp(self, s, position, left_margin=0.0, upper_r=None):
        a, f = self.a, self.f
        if s.coords.shape[1] == 0:
            return
        if upper_r is None:  # make_halfstep / chunk_particles: the solver's limit (chimera_main.py:67-68, 317-318)
            upper_r = a["Rgrid"].max()
        dom = np.asfortranarray([a["leftX"] + left_margin, a["rightX"], 0.0, upper_r ** 2])
        src = s.coords_halfstep if position == "cntr" else s.coords
        if self.chunked:
            ids, s.chunks, go_out = f.chunk_coords_boundaries(src, dom, a["Xgrid"], self.nchnk)
            keep = ids.argsort(kind="stable")[go_out:]
        else:
            keep, n = f.sortpartsout(s.coords, dom)
            keep = keep[:n]
        n = keep.shape[0]
        for k in ("coords", "coords_halfstep", "momenta"):
            v = f.align_data_vec(getattr(s, k), keep)
            setattr(s, k, np.asfortranarray(v[:, :n]).copy(order="F"))
        s.weights = f.align_data_scl(s.weights, keep)[:n].copy()

    # ---- chimera_main.py:153-248 ---------------------------------------------------------------
    def _dep(self, kind, grid, s, coords):
        a, f = self.a, self.f
        name = "dep_" + kind + ("_env" if self.env else "") + ("_chnk" if self.chunked else "")
        args = [coords] + ([s.momenta] if kind == "curr" else []) + [s.weights, grid]
        if self.chunked:
            args += [s.chunks, self.guards]
        return getattr(f, name)(*args, a["leftX"], *a["DepProj"])

    def project_current(self):
        self.J[:] = 0.0
        for s in self.sp:
            if s.still or s.coords.shape[1] == 0:
                continue
            self.J = self._dep("curr", self.J, s, s.coords_halfstep)
        if self.reduce:
            self.reduce(self.J)
        a, f = self.a, self.f
        self.J_fb = f.fb_vec_in(self.J_fb, self.J, a["leftX"], *a["FBCurrIn"])
        self.J_fb = f.omp_mult_vec(self.J_fb, a["DepFact"])

    def dep_bg(self):
        self.Bck[:] = 0.0
        for s in self.sp:
            if s.still and s.coords.shape[1]:
                self.Bck = self._dep("dens", self.Bck, s, s.coords)
        if self.reduce:
            self.reduce(self.Bck)

    def project_density(self):
        if not (self.space_charge or self.static_kick):
            return
        a, f = self.a, self.f
        if not self.static_kick:
            self.g_prv[:] = self.g_nxt
        self.Rho[:] = 0.0
        if self.rank == 0:  # the (already summed) background enters the all-reduced density once
            self.Rho += self.Bck
        for s in self.sp:
            if s.still or s.coords.shape[1] == 0:
                continue
            self.Rho = self._dep("dens", self.Rho, s, s.coords_halfstep if self.static_kick else s.coords)
        if self.reduce:
            self.reduce(self.Rho)
        self.Rho_fb = f.fb_scl_in(self.Rho_fb, self.Rho, a["leftX"], *a["FBCurrIn"])
        self.Rho_fb = f.omp_mult_scl(self.Rho_fb, a["DepFact"])
        grad = f.fb_grad_env if self.env else f.fb_grad
        self.g_nxt = grad(self.g_nxt, self.Rho_fb, *a["FBDiff"])

    # ---- solvers.py:281-331 --------------------------------------------------------------------
    def update_fields(self):
        a, f = self.a, self.f
        graddiv = f.fb_graddiv_env if self.env else f.fb_graddiv
        if self.static_kick:  # chimera_main.py:118-125, solvers.py:333-406
            self.EG_fb[:] = 0.0
            for s in self.sp:
                px = (s.momenta[0] * s.weights).sum() / s.weights.sum()
                beta = px / np.sqrt(1 + px ** 2)
                if self.npoiss:
                    self.vec_fb[:] = self.J_fb
                    self.vec_fb = graddiv(self.vec_fb, *a["FBDiff"])
                    self.J_fb = f.poiss_corr_stat(self.J_fb, self.vec_fb, self.g_nxt, -1j * beta * a["kx"], a["PoissFact"])
                self.maxwell_solver_stat(px)
                self.EG_fb = f.field_drift(self.EG_fb, a["kx"], beta, a["dt"])
            return
        for _ in range(self.npoiss):
            self.vec_fb[:] = self.J_fb
            self.vec_fb = graddiv(self.vec_fb, *a["FBDiff"])
            if self.space_charge:
                self.J_fb = f.poiss_corr(self.J_fb, self.vec_fb, self.g_prv, self.g_nxt, a["dt_inv"], a["PoissFact"])
            else:
                self.vec_fb = f.omp_mult_vec(self.vec_fb, a["PoissFact"])
                self.J_fb = f.omp_add_vec(self.J_fb, self.vec_fb)
        if self.space_charge:
            self.EG_fb = f.maxwell_push_with_spchrg(self.EG_fb, self.J_fb, self.g_prv, self.g_nxt, self.PE, self.PG)
        else:
            self.EG_fb = f.maxwell_push_wo_spchrg(self.EG_fb, self.J_fb, self.PE, self.PG)

    def maxwell_solver_stat(self, px0):  # solvers.py:333-358
        if not (self.space_charge or self.static_kick):
            return
        c1, c2 = self.S.static_coeffs(px0)
        self.EG_fb = self.f.maxwell_init_push(self.EG_fb, self.J_fb, self.g_nxt, c1, c2)

    # ---- chimera_main.py:130-151, solvers.py:450-465, 536-553 ------------------------------------
    def project_fields(self):
        a, f = self.a, self.f
        rot = f.fb_rot_env if self.env else f.fb_rot
        self.B_fb = rot(self.B_fb, self.EG_fb[:, :, :, 3:], *a["FBDiff"])
        self.B_fb = f.omp_mult_vec(self.B_fb, a["PoissFact"])
        self.EB = f.fb_eb_out(self.EB, self.EG_fb, self.B_fb, a["leftX"], *a["FBout"])
        self.EB = (f.eb_correction_env if self.env else f.eb_correction)(self.EB)
        proj = f.proj_fld_env if self.env else f.proj_fld
        for s in self.sp:
            if s.still:
                continue
            # Specie.make_field (species.py:246-256): resize when the particle count changed, then zero in place
            if s.EB.shape[-1] != s.coords.shape[1]:
                s.EB.resize((6, s.coords.shape[1]), refcheck=False)
            s.EB[:] = 0.0
            if s.coords.shape[1] == 0:
                continue
            s.EB = proj(s.coords, s.weights, self.EB, s.EB, a["leftX"], *a["DepProj"])

    def devices_and_push(self, dt_frac):
        a, f = self.a, self.f
        for s in self.sp:
            if s.still or s.coords.shape[1] == 0:
                continue
            for dev in s.devices:  # species.py:274-277
                args = [np.asfortranarray(v, dtype=float) if isinstance(v, (list, tuple, np.ndarray)) else v for v in dev[1:]]
                s.EB = dev[0](s.coords, s.EB, self.istep * a["dt"], *args)
            s.momenta = f.push_velocs(s.momenta, s.EB, s.push_fact * a["dt"] * dt_frac)

    # ---- chimera_main.py:250-304: moving window, stage 1 ----------------------------------------
    def frame_act(self, wind, add=None):
        """damp_fields, move_frame, add_plasma (``add`` = {species index: (coords, momenta, weights)} as produced
        by the driver's gen_parts), damp_plasma, postframe_corr -- in the reference's order"""
        a, f = self.a, self.f
        if wind.get("AbsorbLayer", 0) > 0:  # solvers.py:619-633 damp_field('left', damp_b=False)
            prof = self.S.get_damp_profile(wind["AbsorbLayer"])
            self.EG_fb[:, :, :, :3] = f.fb_filtr(self.EG_fb[:, :, :, :3], a["leftX"], a["kx"], prof, 0)
            self.EG_fb[:, :, :, 3:] = f.fb_filtr(self.EG_fb[:, :, :, 3:], a["leftX"], a["kx"], prof, 0)
        a["Xgrid"] = a["Xgrid"] + wind["shiftX"]
        a["leftX"], a["rightX"] = a["Xgrid"][0], a["Xgrid"][-1]
        for i, (x, p, w) in (add or {}).items():  # species.py:218-244
            s = self.sp[i]
            s.coords = np.asfortranarray(np.concatenate((s.coords, x), axis=1))
            s.coords_halfstep = np.asfortranarray(np.concatenate((s.coords_halfstep, x), axis=1))
            s.momenta = np.asfortranarray(np.concatenate((s.momenta, p), axis=1))
            s.weights = np.concatenate((s.weights, w))
        if "AbsorbLayer" in wind:
            # species.py:373-376: the window culls at the SPECIES' upperR; its r grid (species.py:85-92) has
            # round(lengthR/dr) nodes, one less than the solver's
            upper_r = a["dr"] * ((a["Nkr"] - 1) - 0.5)
            for s in self.sp:
                self.chunk_and_damp(s, "stag", left_margin=wind["AbsorbLayer"] * a["dx"], upper_r=upper_r)
        if self.space_charge:  # postframe_corr
            if self.background:
                self.dep_bg()
            self.Rho[:] = 0.0
            if self.rank == 0:
                self.Rho += self.Bck
            for s in self.sp:
                if s.still or s.coords.shape[1] == 0:
                    continue
                self.Rho = self._dep("dens", self.Rho, s, s.coords)
            if self.reduce:
                self.reduce(self.Rho)

    # ---- chimera_main.py:61-92 -----------------------------------------------------------------
    def make_halfstep(self, px0=(0.0,)):
        for s in self.sp:
            self.chunk_and_damp(s, "stag")
        if self.background:
            self.dep_bg()
        self.project_current()
        self.project_density()
        for p in px0:
            self.maxwell_solver_stat(p)
        self.project_fields()
        self.devices_and_push(0.5)

    def move_frame(self, shift):  # chimera_main.py:286-290
        a = self.a
        a["Xgrid"] = a["Xgrid"] + shift
        a["leftX"], a["rightX"] = a["Xgrid"][0], a["Xgrid"][-1]

    def make_step(self):
        f, a = self.f, self.a
        self.istep += 1
        if any(self.window):
            self.move_frame(self.window[0])  # frame_act(istep): stage 1 (chimera_main.py:83)
        for s in self.sp:
            if s.still or s.coords.shape[1] == 0:
                continue
            s.coords, s.coords_halfstep = f.push_coords(s.coords, s.momenta, s.coords_halfstep, a["dt"])
        if self.sort_every > 0 and self.istep % self.sort_every == 0:
            for s in self.sp:
                self.chunk_and_damp(s, "cntr")
        self.project_current()
        if any(self.window):
            self.move_frame(self.window[1])  # frame_act(istep, 'stage2') of a 'Staged' frame (chimera_main.py:87)
        self.project_density()
        self.update_fields()
        self.project_fields()
        self.devices_and_push(1.0)
