"""CPU stand-in for libchimera_b200's resident engine (TEST INFRASTRUCTURE) so that the HOST-SIDE multi-rank logic of
chimera_b200/engine.py -- the Python that sequences the phases across ranks: ``Engine.make_halfstep``, ``Engine.step``
(fused / unfused schedule, re-binning cadence, per-step window), ``_fields_out`` (E and B halves all-gathered
separately and pipelined), ``_allreduce_grids_async``, the rank-0 background rule -- runs UNCHANGED on a world_size-2
``gloo`` job without a GPU (tests/test_dist_gloo.py).

``CpuEngine`` subclasses ``Engine`` and replaces only what touches the CUDA library: construction, the named arrays
and ``run(phase)``.  Each phase of csrc/engine.cu is executed here by the oracle (particle work, full-grid forward
transform) and by the numpy restatement (slab-local spectral calculus with the mirror shift), following the phase
definitions in include/chimera_b200.h (enum chimera_engine_phase).  Nothing here is imported by the product.
"""
import copy

import numpy as np

from chimera_b200 import sharding
from chimera_b200.engine import Engine, EngineConfig
from oracle import fimera as ofim
from oracle import np_ref
from pic_ref import RefRun, RefSpecies


class CpuEngine(Engine):
    def __init__(self, setup, group=None, slab=None, sort_every=None):  # noqa: super().__init__ would load the CUDA library
        import torch.distributed as dist

        self.setup = setup
        a = self.a = setup.Args
        feats = a.get("Features", ())
        cfg = EngineConfig()
        cfg.env = int(setup.env)
        cfg.space_charge = int("SpaceCharge" in feats)
        cfg.poisson_iters = 0 if "NoPoissonCorrection" in feats else 3
        cfg.chunked = int("Xchunked" in a)
        cfg.nchnk, cfg.guards = (int(a["Xchunked"][0]), int(a["Xchunked"][1])) if cfg.chunked else (1, 0)
        cfg.sort_every = (cfg.guards + 1 if cfg.chunked else 0) if sort_every is None else sort_every
        cfg.nx, cfg.nrn, cfg.nkr, cfg.nm = a["Nx"], a["Nr"], a["Nkr"], a["Mtot"]
        cfg.dt = a["dt"]
        self._dist, self._group, self.group = dist, None, group
        self.rank, self.world = (dist.get_rank(), dist.get_world_size()) if group is not None else (0, 1)
        self.slab_rank, self.slab_world = self.rank, self.world
        if slab is None:
            slab = self.world > 1 and sharding.slab_supported(a["Nx"], self.world)
        self.slab = bool(slab) and self.world > 1
        self.rows = np.arange(a["Nx"])
        self.mirror = 0
        if self.slab:
            self.rows = sharding.kx_slab_rows(a["Nx"], self.rank, self.world)
            cfg.nx_slab = int(self.rows.size)
            self.mirror = cfg.mirror_shift = sharding.mirror_shift(self.rank, self.world)
        self.cfg = cfg
        self.istep, self.fuse, self.overlap, self._win = 0, True, True, (0.0, 0.0)
        self.xgrid, self.leftX_J = a["Xgrid"].copy(), None  # window position (ChimeraRun.move_frame shifts Xgrid)
        self.time = 0.0
        # the oracle-side state holder: particle containers and full grids (its own copy of the solver dictionary is
        # not needed: the window position lives in self.leftX)
        self.r = RefRun(ofim, setup, [], sort_every=0)
        L = self.rows.size
        fb = lambda c: np.zeros((L, a["Nkr"], a["Mtot"]) + ((c,) if c else ()), dtype=complex, order="F")  # noqa: E731
        self.arr = {"J": self.r.J, "Rho": self.r.Rho, "BckGrndRho": self.r.Bck, "EB": self.r.EB,
                    "EG_fb": fb(6), "J_fb": fb(3), "B_fb": fb(3), "Rho_fb": fb(0), "gradRho_fb_prv": fb(3),
                    "gradRho_fb_nxt": fb(3), "CPSATD1": fb(2), "CPSATD2": fb(2),
                    "EB_slab": np.zeros((L, a["Nr"], a["Mtot"], 6), dtype=complex, order="F"),
                    "EB_gath": np.zeros(a["Nx"] * a["Nr"] * a["Mtot"] * 6, dtype=complex)}

    # ---- the named arrays ------------------------------------------------------------------------------------
    def device_tensor(self, name, dtype=np.float64):
        import torch

        flat = self.arr[name].ravel(order="K")
        assert np.shares_memory(flat, self.arr[name])
        return torch.from_numpy(flat.view(np.float64))

    def upload(self, name, arr):
        self.arr[name][...] = arr[self.rows] if (self.slab and name.endswith(("_fb", "PSATD1", "PSATD2"))) else arr

    def download(self, name):
        return self.arr[name].copy(order="F")

    def add_species(self, coords, momenta, weights, charge=-1.0, mass=1.0, still=False, **_):
        self.r.sp.append(RefSpecies(coords, momenta, weights, charge=charge, mass=mass, still=still))
        return len(self.r.sp) - 1

    def particles(self, sid=0):
        s = self.r.sp[sid]
        return s.coords, s.coords_halfstep, s.momenta, s.weights

    def count(self, sid=0):
        return self.r.sp[sid].weights.size

    def set_time(self, t):
        self.time = t

    def set_window(self, velocity, time_step=None, staged=False):
        ts = self.cfg.dt if time_step is None else time_step
        self._win = (0.5 * velocity * ts, 0.5 * velocity * ts) if staged else (velocity * ts, 0.0)

    def sync(self):
        pass

    # ---- moving window: thin wrappers over library calls in the product, restated on the oracle; Engine.frame_act
    # itself (the multi-rank logic under test) is inherited
    def damp_field(self, profile, config="left"):
        import torch

        a, A = self.a, self.arr
        self._sync_dict()
        mode = {"left": 0, "right": 1, "both": 2}[config]
        eg = A["EG_fb"]
        if self.slab:  # gather the slabs of all ranks, rebuild the full rows (chimera_engine_damp_field_slab)
            mine = self.device_tensor("EG_fb")
            gath = torch.zeros(mine.numel() * self.world, dtype=mine.dtype)
            self._dist.all_gather_into_tensor(gath, mine, group=self._group)
            full = np.zeros((a["Nx"], a["Nkr"], a["Mtot"], 6), dtype=complex, order="F")
            n = mine.numel() // 2
            for r in range(self.world):
                blk = gath.numpy().view(np.complex128)[r * n:(r + 1) * n].reshape(eg.shape, order="F")
                full[sharding.kx_slab_rows(a["Nx"], r, self.world)] = blk
            eg = full
        for h in (slice(0, 3), slice(3, 6)):
            eg[..., h] = ofim.fb_filtr(np.asfortranarray(eg[..., h]), self.leftX, a["kx"], np.ascontiguousarray(profile), mode)
        A["EG_fb"][...] = eg[self.rows] if self.slab else eg

    def move_window(self, shift):
        self.xgrid = self.xgrid + shift
        self._sync_dict()

    def append_particles(self, sid, coords, momenta, weights):
        s = self.r.sp[sid]
        s.coords = np.asfortranarray(np.concatenate((s.coords, coords), axis=1))
        s.coords_halfstep = np.asfortranarray(np.concatenate((s.coords_halfstep, coords), axis=1))
        s.momenta = np.asfortranarray(np.concatenate((s.momenta, momenta), axis=1))
        s.weights = np.concatenate((s.weights, weights))

    def sort(self, on_halfstep=False, left_margin=0.0, upper_r=None):
        self._sync_dict()
        for s in self.r.sp:
            self.r.chunk_and_damp(s, "cntr" if on_halfstep else "stag", left_margin=left_margin, upper_r=upper_r)

    def close(self):
        pass

    # ---- phases (include/chimera_b200.h enum chimera_engine_phase; csrc/engine.cu ph_*) --------------------------
    @property
    def leftX(self):
        return self.xgrid[0]

    def _sync_dict(self):  # the oracle helpers read the window position from the solver dictionary
        self.a["Xgrid"], self.a["leftX"], self.a["rightX"] = self.xgrid, self.xgrid[0], self.xgrid[-1]

    def run(self, phase, arg=0.0):
        self._sync_dict()
        getattr(self, "_ph_" + phase)(arg)

    def _moving(self):
        return [s for s in self.r.sp if not s.still and s.coords.shape[1]]

    def _ph_window(self, arg):
        self.xgrid = self.xgrid + self._win[1 if arg == 2.0 else 0]
        self._sync_dict()

    def _ph_sort(self, arg):
        for s in self.r.sp:
            self.r.chunk_and_damp(s, "cntr" if arg else "stag")

    def _ph_push_coords(self, arg):
        for s in self._moving():
            s.coords, s.coords_halfstep = ofim.push_coords(s.coords, s.momenta, s.coords_halfstep, self.a["dt"])

    def _ph_deposit_J(self, arg):
        r = self.r
        r.J[:] = 0.0
        for s in self._moving():
            r.J = self.arr["J"] = r._dep("curr", r.J, s, s.coords_halfstep)
        self.leftX_J = self.leftX

    def _ph_deposit_rho(self, arg):
        r = self.r
        r.Rho[:] = 0.0
        if arg:
            r.Rho += r.Bck
        for s in self._moving():
            r.Rho = self.arr["Rho"] = r._dep("dens", r.Rho, s, s.coords)

    def _ph_deposit_bg(self, arg):
        r = self.r
        r.Bck[:] = 0.0
        for s in r.sp:
            if s.still and s.coords.shape[1]:
                r.Bck = self.arr["BckGrndRho"] = r._dep("dens", r.Bck, s, s.coords)

    def _ph_fb_in_J(self, arg):
        a = self.a
        lx = self.leftX if self.leftX_J is None else self.leftX_J
        self.leftX_J = None
        full = ofim.omp_mult_vec(ofim.fb_vec_in(self.setup.zeros_fb(3), self.r.J, lx, *a["FBCurrIn"]), a["DepFact"])
        self.arr["J_fb"][...] = full[self.rows]

    def _calc(self):  # slab-local spectral calculus: mirror partner inside the slab (chimera_b200/sharding.py)
        np_ref.MIRROR_SHIFT = self.mirror
        Dp, Dm, kx = self.a["FBDiff"]
        return Dp, Dm, kx[self.rows]

    def _ph_fb_in_rho(self, arg):
        a, A = self.a, self.arr
        A["gradRho_fb_prv"][...] = A["gradRho_fb_nxt"]
        full = ofim.omp_mult_scl(ofim.fb_scl_in(self.setup.zeros_fb(), self.r.Rho, self.leftX, *a["FBCurrIn"]), a["DepFact"])
        A["Rho_fb"][...] = full[self.rows]
        Dp, Dm, kx = self._calc()
        A["gradRho_fb_nxt"][...] = (np_ref.fb_grad_env if self.cfg.env else np_ref.fb_grad)(A["Rho_fb"], Dp, Dm, kx)
        np_ref.MIRROR_SHIFT = 0

    def _ph_init_push(self, arg):
        A = self.arr
        A["EG_fb"][...] = np_ref.maxwell_init_push(A["EG_fb"], A["J_fb"], A["gradRho_fb_nxt"], A["CPSATD1"], A["CPSATD2"])

    def _ph_poisson(self, arg):
        a, A = self.a, self.arr
        Dp, Dm, kx = self._calc()
        j = A["J_fb"]
        graddiv = np_ref.fb_graddiv_env if self.cfg.env else np_ref.fb_graddiv
        for _ in range(self.cfg.poisson_iters):
            gd = graddiv(j, Dp, Dm, kx)
            if self.cfg.space_charge:
                j = np_ref.poiss_corr(j, gd, A["gradRho_fb_prv"], A["gradRho_fb_nxt"], a["dt_inv"], a["PoissFact"][self.rows])
            else:  # solvers.py:322-326: J += PoissFact grad div J
                j = j + gd * a["PoissFact"][self.rows][..., None]
        A["J_fb"][...] = j
        np_ref.MIRROR_SHIFT = 0

    def _ph_maxwell(self, arg):
        A, S = self.arr, self.setup
        if self.cfg.space_charge:
            A["EG_fb"][...] = np_ref.maxwell_push_with_spchrg(A["EG_fb"], A["J_fb"], A["gradRho_fb_prv"], A["gradRho_fb_nxt"],
                                                             S.PSATD_E[self.rows], S.PSATD_G[self.rows])
        else:
            A["EG_fb"][...] = np_ref.maxwell_push_wo_spchrg(A["EG_fb"], A["J_fb"], S.PSATD_E[self.rows], S.PSATD_G[self.rows])

    def _backward_slab(self, src):
        """backward DHT + phase of the slab rows, no inverse x-FFT (fb_out_slab_dev): (L, Nkr, M, 3) -> (L, Nr, M, 3)"""
        a = self.a
        kxo, Out = a["FBout"]
        out = np.zeros((self.rows.size, a["Nr"], a["Mtot"], 3), dtype=complex)
        for m in range(a["Mtot"]):
            out[:, 1:, m, :] = np.einsum("xkc,kr->xrc", src[:, :, m, :], Out[:, :, m])
        return out * np.exp(1j * kxo[self.rows] * self.leftX)[:, None, None, None]

    def _ph_fields_out_a(self, arg):
        a, A = self.a, self.arr
        part = 1 if arg == 1.0 else (2 if arg == 2.0 else 0)
        if part != 1:
            Dp, Dm, kx = self._calc()
            rot = np_ref.fb_rot_env if self.cfg.env else np_ref.fb_rot
            A["B_fb"][...] = rot(A["EG_fb"][..., 3:], Dp, Dm, kx) * a["PoissFact"][self.rows][..., None]
            np_ref.MIRROR_SHIFT = 0
        if part != 2:
            A["EB_slab"][..., :3] = self._backward_slab(A["EG_fb"][..., :3])
        if part != 1:
            A["EB_slab"][..., 3:] = self._backward_slab(A["B_fb"])

    def _ph_fields_out_b(self, arg):
        a, A = self.a, self.arr
        part = 1 if arg == 1.0 else (2 if arg == 2.0 else 0)
        L, nx = self.rows.size, a["Nx"]
        world = nx // L
        shp3 = (L, a["Nr"], a["Mtot"], 3)
        n3 = int(np.prod(shp3))
        for h in ((0, 1) if part == 0 else (part - 1,)):
            full = np.zeros((nx, a["Nr"], a["Mtot"], 3), dtype=complex)
            for r in range(world):
                rr = sharding.kx_slab_rows(nx, r, world)
                if part == 0:  # [rank][(L, Nr, M, 6)]
                    blk = A["EB_gath"][r * 2 * n3:(r + 1) * 2 * n3].reshape((L, a["Nr"], a["Mtot"], 6), order="F")[..., 3 * h:3 * h + 3]
                else:          # [half][rank][(L, Nr, M, 3)]
                    blk = A["EB_gath"][(h * world + r) * n3:(h * world + r + 1) * n3].reshape(shp3, order="F")
                full[rr] = blk
            eb = np.fft.ifft(full, axis=0) * nx  # unnormalised backward FFT (Q9)
            if self.cfg.env:  # eb_correction_env on this half (grid_deps_env.f90:240-283)
                eb /= np.pi
                eb[:, 0] = eb[:, 1] if a["Mtot"] == 1 else -eb[:, 1]
            else:             # eb_correction (grid_deps.f90:219-266)
                eb[:, :, 0] /= 2 * np.pi
                eb[:, :, 1:] /= np.pi
                eb[:, 0, 0] = eb[:, 1, 0]
                eb[:, 0, 1:] = -eb[:, 1, 1:]
            self.r.EB[..., 3 * h:3 * h + 3] = eb

    def _ph_gather_push(self, arg):
        a, r = self.a, self.r
        for s in self._moving():
            proj = ofim.proj_fld_env if self.cfg.env else ofim.proj_fld
            s.EB = proj(s.coords, s.weights, r.EB, np.zeros((6, s.coords.shape[1]), order="F"), self.leftX, *a["DepProj"])
            s.momenta = ofim.push_velocs(s.momenta, s.EB, s.push_fact * a["dt"] * arg)

    def _ph_particles_fused(self, arg):
        self._ph_gather_push(1.0)
        win = any(self._win)
        if win:
            self._ph_window(1.0)
        self._ph_push_coords(0.0)
        self._ph_deposit_J(0.0)
        if win:
            self._ph_window(2.0)
        if self.cfg.space_charge:
            self._ph_deposit_rho(arg)

    def _ph_gather_push_coords(self, arg):  # CHB_GATHER_PUSH_COORDS: before a re-binning step's sort; window stage 1
        self._ph_gather_push(1.0)
        self._ph_window(1.0)
        self._ph_push_coords(0.0)

    def _ph_deposit_fused(self, arg):  # CHB_DEPOSIT_FUSED: after the sort; window stage 2 between J and rho
        self._ph_deposit_J(0.0)
        self._ph_window(2.0)
        if self.cfg.space_charge:
            self._ph_deposit_rho(arg)


def reference_run(setup, species, eg0, nsteps, window=(0.0, 0.0)):
    """the single-process reference sequence the ranks are compared with"""
    S = copy.copy(setup)
    S.Args = copy.deepcopy(setup.Args)
    ref = RefRun(ofim, S, [RefSpecies(*sp[:3], **sp[3]) for sp in species], background=any(sp[3].get("still") for sp in species))
    ref.EG_fb[:] = eg0
    ref.window = window
    ref.make_halfstep(px0=(0.0,) * len(species))
    for _ in range(nsteps):
        ref.make_step()
    return ref
