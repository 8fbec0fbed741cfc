"""Device-resident PIC engine: host-side mirror of the reference's ``ChimeraRun`` step loop.

The reference sequences one time step in Python (``ChimeraRun.make_step``, reference
moduls/chimera_main.py:82-92; ``make_halfstep`` :61-80) and crosses into Fortran ~21 times per
step with host arrays.  Here the same sequence runs inside libchimera_b200.so on arrays that stay in
HBM (csrc/engine.cu); this module only builds the configuration from a :class:`SolverSetup`, uploads
the operator tables once and exposes the state.

Multi-GPU (one process per GPU, ``torch.distributed``): particles are sharded across ranks, every
rank deposits into its own J / Rho grids, the grids are summed with an NCCL all-reduce over NVLink and
the spectral solve is sharded by kx slab (mirror pairs of rows, chimera_b200/sharding.py): every rank
x-FFTs the summed grids, transforms / corrects / advances only its rows, and the backward-transformed
slabs are all-gathered before the inverse x-FFT and the gather to the rank's own particles.  The two
collectives are pipelined with their neighbours (``Engine.overlap``): the E half of the slab is gathered while
the B half is computed, the reduction of Rho runs behind the forward transform of J.  There is
no CPU fallback: without the CUDA library the import of :mod:`chimera_b200._lib` fails.

Also mirrored here: the reference's moving frames (``frame_act`` for windows that act every few steps with
``AbsorbLayer`` / ``AddPlasma``; ``set_window`` for frames that move every step, 'Staged' or not), per-species
external-field devices, the 'StaticKick' schedule and the integrated diagnostics of moduls/diagnostics.py.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, sharding

_i64 = ctypes.c_longlong

PHASES = (
    "push_coords", "sort", "deposit_J", "deposit_rho", "deposit_bg", "fb_in_J", "fb_in_rho", "poisson",
    "maxwell", "init_push", "fields_out", "gather_push", "add_bg", "fields_out_a", "fields_out_b",
    "particles_fused", "static_fields", "window", "gather_push_coords", "deposit_fused",
    "col_fwd", "fb_in_col", "col_bwd", "eb_finish",
)
PHASE_ID = {n: i for i, n in enumerate(PHASES)}


class EngineConfig(ctypes.Structure):
    """Mirror of ``chimera_engine_config`` (include/chimera_b200.h)."""

    _fields_ = [
        ("env", ctypes.c_int), ("space_charge", ctypes.c_int), ("poisson_iters", ctypes.c_int),
        ("coef_complex", ctypes.c_int), ("chunked", ctypes.c_int), ("nchnk", ctypes.c_int),
        ("guards", ctypes.c_int), ("sort_every", ctypes.c_int), ("undulator", ctypes.c_int),
        ("nx", _i64), ("nrn", _i64), ("nkr", _i64), ("nm", _i64),
        ("leftX", ctypes.c_double), ("rightX", ctypes.c_double), ("dx", ctypes.c_double),
        ("dr", ctypes.c_double), ("dt", ctypes.c_double), ("kx0", ctypes.c_double),
        ("rcull2", ctypes.c_double), ("chunk_len", ctypes.c_double),
        ("und_a0", ctypes.c_double), ("und_lambda", ctypes.c_double), ("und_X0", ctypes.c_double),
        ("und_Lx", ctypes.c_double), ("nx_slab", _i64), ("mirror_shift", ctypes.c_int), ("static_kick", ctypes.c_int),
    ]


class EngineError(RuntimeError):
    pass


class _DevView:
    """Zero-copy handle on an engine array for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr, nbytes, dtype):
        n = nbytes // np.dtype(dtype).itemsize
        self.__cuda_array_interface__ = {
            "shape": (n,), "typestr": np.dtype(dtype).str, "data": (ptr, False), "version": 2, "strides": None,
        }


class Engine:
    """One solver + its particle species on one GPU.

    Parameters
    ----------
    setup : SolverSetup
        grids, operators and PSATD tables (chimera_b200.solver_setup, mirrors solvers.py:27-279)
    chunked : bool, optional
        use the ``*_chnk`` deposition semantics (default: the solver dict has ``Xchunked``)
    sort_every : int, optional
        re-binning cadence in steps (reference: ``Xchunked[1]+1``, chimera_main.py:310); 0 = never
    poisson_iters : int, optional
        reference default 3 (solvers.py:301); forced to 0 by the ``NoPoissonCorrection`` feature
    undulator : dict, optional
        ``{'a0','lambda','X0','Lx'}`` of ``undul_analytic`` (devices.f90:162)
    group : torch.distributed process group or True, optional
        shard particles over the ranks of the group and all-reduce the deposited grids
    slab : bool or (rank, world), optional
        shard the spectral solve by kx slab (chimera_b200.sharding): this engine then holds ``Nx/world``
        rows of every spectral array.  Default: on under a process group whenever ``Nx`` is divisible by
        ``2*world``.  ``(rank, world)`` selects a slab without a process group (tests drive the exchange).
    """

    colflow = False       # column-block dataflow of the multi-rank solve in use (_setup_colflow)
    comm_profile = False  # time every collective on its own (comm_timings)

    def __init__(self, setup, chunked=None, sort_every=None, poisson_iters=None, undulator=None, group=None, slab=None):
        self.lib = _lib.load()
        self.setup = setup
        a = setup.Args
        feats = a.get("Features", ())
        cfg = EngineConfig()
        cfg.static_kick = int("StaticKick" in feats)
        cfg.env = int(setup.env)
        cfg.space_charge = int("SpaceCharge" in feats)
        if poisson_iters is None:
            poisson_iters = 0 if "NoPoissonCorrection" in feats else 3
        cfg.poisson_iters = int(poisson_iters)
        cfg.coef_complex = int(np.iscomplexobj(setup.PSATD_E))
        if chunked is None:
            chunked = "Xchunked" in a
        cfg.chunked = int(bool(chunked))
        cfg.nchnk, cfg.guards = (int(a["Xchunked"][0]), int(a["Xchunked"][1])) if chunked else (1, 0)
        if sort_every is None:
            sort_every = cfg.guards + 1 if chunked else 0
        cfg.sort_every = int(sort_every)
        cfg.nx, cfg.nrn, cfg.nkr, cfg.nm = a["Nx"], a["Nr"], a["Nkr"], a["Mtot"]
        cfg.leftX, cfg.rightX, cfg.dx, cfg.dr, cfg.dt, cfg.kx0 = a["leftX"], a["rightX"], a["dx"], a["dr"], a["dt"], a["kx0"]
        cfg.rcull2 = float(a["Rgrid"].max() ** 2)
        xg = a["Xgrid"]
        cfg.chunk_len = float(xg[a["Nx"] // cfg.nchnk] - xg[0]) if cfg.nchnk > 1 else float(xg[-1] - xg[0])
        if undulator:
            cfg.undulator = 1
            cfg.und_a0, cfg.und_lambda, cfg.und_X0, cfg.und_Lx = (float(undulator[k]) for k in ("a0", "lambda", "X0", "Lx"))
        self.group = group
        self.rank, self.world = 0, 1
        if group is not None:
            import torch.distributed as dist

            self._dist = dist
            self._group = None if group is True else group
            self.rank, self.world = dist.get_rank(self._group), dist.get_world_size(self._group)
        # kx-slab sharding of the spectral solve
        self.slab_rank, self.slab_world = self.rank, self.world
        if isinstance(slab, tuple):
            self.slab_rank, self.slab_world = int(slab[0]), int(slab[1])
            slab = True
        if slab is None:
            slab = self.slab_world > 1 and sharding.slab_supported(a["Nx"], self.slab_world)
        self.slab = bool(slab) and self.slab_world > 1
        self.rows = None
        if self.slab:
            self.rows = sharding.kx_slab_rows(a["Nx"], self.slab_rank, self.slab_world)
            cfg.nx_slab = int(self.rows.size)
            cfg.mirror_shift = sharding.mirror_shift(self.slab_rank, self.slab_world)
        self.cfg = cfg
        self._h = ctypes.c_void_p()
        self._check(self.lib.chimera_engine_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        kx_base = a["FBCurrIn"][0]
        for name, arr in (
            ("InCurr", a["InCurr"]), ("Out", a["Out"]), ("DpS2S", a["DpS2S"]), ("DmS2S", a["DmS2S"]),
            ("kx", a["kx"]), ("kx_base", kx_base), ("DepFact", a["DepFact"]), ("PoissFact", a["PoissFact"]),
            ("PSATD_E", setup.PSATD_E), ("PSATD_G", setup.PSATD_G), ("Rgrid", a["Rgrid"]),
        ) + ((("w", a["w"]),) if cfg.static_kick else ()):
            self.upload(name, arr)
        if self.slab:
            gmap = np.empty(a["Nx"], dtype=np.int64)
            for r in range(self.slab_world):
                rr = sharding.kx_slab_rows(a["Nx"], r, self.slab_world)
                gmap[rr] = r * rr.size + np.arange(rr.size)
            self._upload_raw("slab_rows", self.rows.astype(np.int64))
            self._upload_raw("gather_map", gmap)
        self.nspecies = 0
        self._px0 = []
        self.istep = 0
        self._pinned = []
        self.fuse = True
        self.overlap = True  # pipeline the collectives with the neighbouring kernels (multi-rank only)
        if group is not None:
            import torch

            # run on torch's current stream so that the NCCL collectives are ordered with the kernels
            self.use_stream(torch.cuda.current_stream().cuda_stream)
        self.colflow = False
        self._comm_events = []
        if group is not None:
            self._setup_colflow()

    # -- plumbing ----------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise EngineError("libchimera_b200: status %d: %s" % (rc, self.lib.chimera_last_error().decode()))

    def close(self):
        if getattr(self, "_pinned", None):
            self.unpin_all()
        if getattr(self, "_h", None):
            self.lib.chimera_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    _KX_FIRST = ("w", "kx", "kx_base", "DepFact", "PoissFact", "PSATD_E", "PSATD_G", "CPSATD1", "CPSATD2", "EG_fb", "J_fb",
                 "B_fb", "Rho_fb", "gradRho_fb_prv", "gradRho_fb_nxt", "vec_fb")

    def _upload_raw(self, name, arr):
        arr = np.asfortranarray(arr)
        self._check(self.lib.chimera_engine_upload(self._h, name.encode(), ctypes.c_void_p(arr.ctypes.data), _i64(arr.nbytes)))

    def upload(self, name, arr):
        """Upload a named array.  On a kx-slab engine, arrays whose first axis is kx may be given in full:
        this rank's rows are taken (chimera_b200.sharding.kx_slab_rows)."""
        arr = np.asarray(arr)
        if self.slab and name in self._KX_FIRST and arr.shape[0] == self.cfg.nx:
            arr = arr[self.rows]
        self._upload_raw(name, arr)

    def shape_of(self, name):
        c = self.cfg
        nxf = c.nx_slab if self.slab else c.nx
        g, f = (c.nx, c.nrn, c.nm), (nxf, c.nkr, c.nm)
        table = {"J": g + (3,), "Rho": g, "BckGrndRho": g, "EB": g + (6,), "EG_fb": f + (6,), "J_fb": f + (3,),
                 "B_fb": f + (3,), "Rho_fb": f, "gradRho_fb_prv": f + (3,), "gradRho_fb_nxt": f + (3,), "vec_fb": f + (3,)}
        return table[name]

    def download(self, name):
        out = np.zeros(self.shape_of(name), dtype=complex, order="F")
        self._check(self.lib.chimera_engine_download(self._h, name.encode(), ctypes.c_void_p(out.ctypes.data), _i64(out.nbytes)))
        return out

    def device_tensor(self, name, dtype=np.float64, raw=False):
        """torch view (no copy) of a named engine array, e.g. for ``torch.distributed.all_reduce``.  J, Rho, EB and EB_slab
        are allocated with spare columns for the column-block dataflow: the view covers the array proper unless ``raw``."""
        import torch

        ptr, nb = ctypes.c_void_p(), _i64()
        self._check(self.lib.chimera_engine_array(self._h, name.encode(), ctypes.byref(ptr), ctypes.byref(nb)))
        t = torch.as_tensor(_DevView(ptr.value, nb.value, dtype), device="cuda")
        if not raw and name in ("J", "Rho", "EB", "EB_slab"):
            c = self.cfg
            shape = self.shape_of(name) if name != "EB_slab" else (c.nx_slab, c.nrn, c.nm, 6)
            t = t[:int(np.prod(shape)) * 16 // np.dtype(dtype).itemsize]
        return t

    # -- particles ---------------------------------------------------------------------------
    def add_species(self, coords, momenta, weights, charge=-1.0, mass=1.0, still=False, coords_half=None, capacity=0,
                    px0=0.0):
        """Add a species; arrays as in ``Specie.Data`` (species.py:122-126); ``px0`` = the species'
        ``MomentaMeans[0]`` (species.py:60), used by ``make_halfstep``'s static kick.  Returns its id."""
        coords = np.asfortranarray(coords, dtype=float)
        momenta = np.asfortranarray(momenta, dtype=float)
        weights = np.asfortranarray(weights, dtype=float)
        ch = None if coords_half is None else np.asfortranarray(coords_half, dtype=float)
        n = coords.shape[1]
        assert coords.shape == (3, n) and momenta.shape == (3, n) and weights.shape == (n,)
        sid = ctypes.c_int(-1)
        self._check(self.lib.chimera_engine_add_species(
            self._h, ctypes.c_void_p(coords.ctypes.data), ctypes.c_void_p(ch.ctypes.data if ch is not None else None),
            ctypes.c_void_p(momenta.ctypes.data), ctypes.c_void_p(weights.ctypes.data), _i64(n),
            ctypes.c_double(2 * np.pi * charge / mass), int(bool(still)), _i64(capacity), ctypes.byref(sid)))
        self.nspecies += 1
        return sid.value

    def add_species_device(self, coords_ptr, momenta_ptr, weights_ptr, n, charge=-1.0, mass=1.0, still=False):
        """Same, from device pointers to (3,n) Fortran-ordered arrays (synthetic benchmarks)."""
        sid = ctypes.c_int(-1)
        self._check(self.lib.chimera_engine_add_species(
            self._h, ctypes.c_void_p(coords_ptr), ctypes.c_void_p(None), ctypes.c_void_p(momenta_ptr),
            ctypes.c_void_p(weights_ptr), _i64(n), ctypes.c_double(2 * np.pi * charge / mass), int(bool(still)),
            _i64(0), ctypes.byref(sid)))
        self.nspecies += 1
        self._px0.append(0.0)
        return sid.value

    DEVICE_KINDS = {"undul_analytic": (1, 4), "undul_analytic_taper": (2, 5), "undul_mapped": (3, 3),
                    "undul_mapped_tap": (4, 5), "planewave": (5, 7), "gaussbeam": (6, 8)}

    def add_device(self, kind, params, a0=0.0, a0_map=None, sid=-1):
        """External-field device of a species, evaluated between the field gather and the momentum push
        (``Specie.make_device``, species.py:258-277; routines of f90/devices.f90).  ``kind`` is the Fortran
        routine name in lower case, ``params`` its ``params`` array; ``a0`` (gaussbeam) and ``a0_map`` (the
        ``a0(2,nx)`` table of the mapped undulators) as in the reference signatures.  ``sid=-1``: every
        non-still species."""
        k, npar = self.DEVICE_KINDS[kind]
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.shape != (npar,):
            raise ValueError("%s takes %d parameters" % (kind, npar))
        mp, nx = None, 0
        if a0_map is not None:
            a0_map = np.asfortranarray(a0_map, dtype=np.float64)
            mp, nx = ctypes.c_void_p(a0_map.ctypes.data), a0_map.shape[1]
        self._check(self.lib.chimera_engine_add_device(self._h, int(sid), k, ctypes.c_double(a0),
                                                       ctypes.c_void_p(params.ctypes.data), npar, mp, _i64(nx)))

    def set_time(self, t):
        """time seen by time-dependent devices in phases driven through ``run`` (``i_step * TimeStep``)"""
        self._check(self.lib.chimera_engine_set_time(self._h, ctypes.c_double(t)))

    # -- moving window (chimera_main.py:250-304) ---------------------------------------------
    def damp_field(self, profile, config="left"):
        """``Solver.damp_field`` (solvers.py:619): absorbing-layer window on E and G in x-space."""
        prof = np.ascontiguousarray(profile, dtype=np.float64)
        mode = {"left": 0, "right": 1, "both": 2}[config]
        if not self.slab:
            self._check(self.lib.chimera_engine_damp_field(self._h, ctypes.c_void_p(prof.ctypes.data), _i64(prof.shape[0]), mode))
            return
        # kx-slab engine: the x-space window needs every kx row of a column -> gather the slabs, filter, keep own rows
        self.damp_field_gather()
        self._check(self.lib.chimera_engine_damp_field_slab(self._h, ctypes.c_void_p(prof.ctypes.data), _i64(prof.shape[0]), mode))

    def damp_field_gather(self):
        """first half of ``damp_field`` on a kx-slab engine: all-gather the ``EG_fb`` slabs into ``EG_gath`` (single-process
        slab emulations fill ``EG_gath`` themselves and call ``chimera_engine_damp_field_slab`` through ``damp_field``)"""
        if not getattr(self, "_damp_ready", False):
            self._check(self.lib.chimera_engine_damp_prepare(self._h))
            self._upload_raw("kx_full", np.asarray(self.setup.Args["kx"], dtype=np.float64))
            self._damp_ready = True
        if self.world > 1:
            self._all_gather_into(self.device_tensor("EG_gath"), self.device_tensor("EG_fb"))

    def set_window(self, velocity, time_step=None, staged=False):
        """A window that moves EVERY step (``MovingFrames`` entry with ``'Steps': 1``; the FEL runs,
        doc/tests/fel-testrun.py:61-63), handled inside ``step``: shifts as ``ChimeraRun.init_Moving_Frames``
        computes them (chimera_main.py:40-51) -- ``'Staged'``: ``Velocity*TimeStep/2`` before ``push_coords`` and again
        between ``dep_curr`` and ``dep_dens``; otherwise ``Velocity*TimeStep`` once, before ``push_coords``.  The fused
        particle kernel deposits on the moved grids, so multi-step calls stay fused.  ``velocity=0`` switches it off."""
        ts = self.cfg.dt if time_step is None else time_step
        if staged:
            s1 = s2 = 0.5 * velocity * ts * 1
        else:
            s1, s2 = velocity * ts * 1, 0.0
        self._win = (s1, s2)
        self._check(self.lib.chimera_engine_set_window(self._h, ctypes.c_double(s1), ctypes.c_double(s2)))

    def move_window(self, shift):
        """``ChimeraRun.move_frame`` (chimera_main.py:286): the engine's grid moves by ``shift``.  The solver
        dictionary ``setup.Args`` belongs to the driver, whose own ``move_frame`` updates it; the engine tracks
        its window in ``self.cfg.leftX / rightX``."""
        self._check(self.lib.chimera_engine_move_window(self._h, ctypes.c_double(shift)))
        self.cfg.leftX += shift
        self.cfg.rightX += shift

    def append_particles(self, sid, coords, momenta, weights):
        """``Specie.add_particles`` (species.py:218): (3,n) coords / momenta and weights(n) join species ``sid``."""
        x = np.asfortranarray(coords, dtype=np.float64)
        p = np.asfortranarray(momenta, dtype=np.float64)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        n = x.shape[1]
        assert x.shape == (3, n) and p.shape == (3, n) and w.shape == (n,)
        self._check(self.lib.chimera_engine_append_particles(self._h, int(sid), ctypes.c_void_p(x.ctypes.data),
                                                             ctypes.c_void_p(p.ctypes.data), ctypes.c_void_p(w.ctypes.data), _i64(n)))

    def sort(self, on_halfstep=False, left_margin=0.0, upper_r=None):
        """``Specie.chunk_and_damp`` (species.py:351) with the absorbing layer ``left_margin`` (in x units);
        ``upper_r``: radial limit of this call (default: the solver's ``Rgrid.max()``, chimera_main.py:318)."""
        if upper_r is None:
            self._check(self.lib.chimera_engine_sort(self._h, int(bool(on_halfstep)), ctypes.c_double(left_margin)))
        else:
            self._check(self.lib.chimera_engine_sort_window(self._h, int(bool(on_halfstep)), ctypes.c_double(left_margin),
                                                            ctypes.c_double(upper_r ** 2)))

    def species_upper_r(self):
        """``Specie.Args['upperR']`` of a species built on the solver's ``Grid`` (species.py:85-92): its r grid has
        ``round(lengthR/dr)`` nodes, one less than the solver's, so a window's ``damp_plasma`` culls one cell earlier"""
        a = self.setup.Args
        return a["dr"] * ((a["Nkr"] - 1) - 0.5)

    def frame_act(self, wind, add=None, background=False):
        """Stage 1 of ``ChimeraRun.frame_act`` (chimera_main.py:292-302) for one moving window ``wind`` (the
        reference's dictionary: ``shiftX``, ``AbsorbLayer`` in cells, ...): damp the fields in the absorbing layer,
        move the grid, add the particles ``add = {species id: (coords, momenta, weights)}`` produced by the
        driver's ``gen_parts`` (host-side: numpy RNG and the user's density profile), cull + re-bin, and redo the
        background / charge density (``postframe_corr``).  Call it between two ``step`` calls."""
        a = self.setup.Args
        if wind.get("AbsorbLayer", 0) > 0:
            self.damp_field(self.setup.get_damp_profile(wind["AbsorbLayer"]))
        self.move_window(wind["shiftX"])
        for sid, (x, p, w) in (add or {}).items():
            self.append_particles(sid, x, p, w)
        if "AbsorbLayer" in wind:
            self.sort(False, wind["AbsorbLayer"] * a["dx"], upper_r=self.species_upper_r())
        if self.cfg.space_charge:
            if background:
                self.deposit_background()
            self.run("deposit_rho", 1.0 if self.rank == 0 else 0.0)
            if self.world > 1:
                self._dist.all_reduce(self.device_tensor("Rho"), group=self._group)

    # -- integrated diagnostics on the device (moduls/diagnostics.py) ---------------------------
    def nrg_out(self):
        """``Diagnostics.nrg_out`` (diagnostics.py:109-124): field energy per kx, rolled as the reference does.
        On a kx-slab engine every rank returns the energies of its own rows (``self.rows``), unrolled."""
        a = self.setup.Args
        nloc = self.cfg.nx_slab if self.slab else a["Nx"]
        out = np.zeros(nloc)
        fact = None
        if not getattr(self, "_energy_fact_up", False):
            ef = np.asfortranarray(a["EnergyFact"][self.rows] if self.slab else a["EnergyFact"], dtype=np.float64)
            fact = ctypes.c_void_p(ef.ctypes.data)
        self._check(self.lib.chimera_engine_field_energy(self._h, fact, ctypes.c_void_p(out.ctypes.data)))
        self._energy_fact_up = True
        if self.slab:
            return out
        return np.r_[out[nloc // 2 + 1:], out[:nloc // 2 + 1]]

    def beam_moments(self, sid=0):
        out = np.zeros(16)
        self._check(self.lib.chimera_engine_beam_moments(self._h, int(sid), ctypes.c_void_p(out.ctypes.data)))
        if self.world > 1:
            import torch

            t = torch.from_numpy(out).cuda()
            self._dist.all_reduce(t, group=self._group)
            out = t.cpu().numpy()
        return out

    def get_beam_envelops(self, sid=0):
        """``Diagnostics.get_beam_envelops`` (diagnostics.py:174-207): centroid, rms size and emittance per axis,
        from 16 sums reduced on the device."""
        m = self.beam_moments(sid)
        sw = m[0]
        xyz0, rms, emit = [], [], []
        for c in range(3):
            wx, wx2, wp2, wxp = m[1 + 5 * c:5 + 5 * c]
            xyz0.append(wx / sw)
            rms.append(np.sqrt(wx2 / sw - (wx / sw) ** 2))
            emit.append(np.sqrt(wx2 / sw * wp2 / sw - wxp ** 2 / sw ** 2))
        return np.array([xyz0, rms, emit])

    def spectrum(self, lo, hi, nbins, quantity="gamma", sid=0):
        """weighted histogram of ``gamma`` or ``px`` of a species: ``np.histogram(q, nbins, (lo, hi), weights=w)[0]``"""
        out = np.zeros(int(nbins))
        self._check(self.lib.chimera_engine_spectrum(self._h, int(sid), {"gamma": 0, "px": 1}[quantity], ctypes.c_double(lo),
                                                     ctypes.c_double(hi), _i64(nbins), ctypes.c_void_p(out.ctypes.data)))
        return out

    def lineout(self, name, ir=0, m=0, l=0):
        """``A[:, ir, m, l]`` of a named complex array, e.g. ``lineout('EB', 0, 0, 0)`` = the on-axis wake field"""
        nx = self.cfg.nx_slab if (self.slab and name not in ("J", "Rho", "BckGrndRho", "EB")) else self.setup.Args["Nx"]
        out = np.zeros(nx, dtype=np.complex128)
        self._check(self.lib.chimera_engine_lineout(self._h, name.encode(), _i64(ir), _i64(m), _i64(l),
                                                    ctypes.c_void_p(out.ctypes.data)))
        return out

    def count(self, sid=0):
        n = _i64()
        self._check(self.lib.chimera_engine_species_count(self._h, sid, ctypes.byref(n)))
        return n.value

    def particles(self, sid=0):
        """(coords, coords_halfstep, momenta, weights) of a species, copied to the host."""
        n = self.count(sid)
        x, xh, p = (np.zeros((3, n), order="F") for _ in range(3))
        w = np.zeros(n)
        self._check(self.lib.chimera_engine_get_species(
            self._h, sid, ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(xh.ctypes.data),
            ctypes.c_void_p(p.ctypes.data), ctypes.c_void_p(w.ctypes.data)))
        return x, xh, p, w

    def chunks(self, sid=0):
        ind = np.zeros(self.cfg.nchnk + 1, dtype=np.int32)
        self._check(self.lib.chimera_engine_get_chunks(self._h, sid, ctypes.c_void_p(ind.ctypes.data)))
        return ind

    # -- stepping ----------------------------------------------------------------------------
    def run(self, phase, arg=0.0):
        self._check(self.lib.chimera_engine_run(self._h, PHASE_ID[phase], ctypes.c_double(arg)))

    def sync(self):
        self._check(self.lib.chimera_engine_sync(self._h))

    def set_fuse(self, on=True):
        """Fuse the particle work between two field solves into one kernel inside multi-step calls (default)."""
        self.fuse = bool(on)
        self._check(self.lib.chimera_engine_set_fuse(self._h, int(self.fuse)))

    def set_graph(self, on=True):
        """Replay the fused step as a CUDA graph between two re-binnings (default on; single GPU, no per-step window)."""
        self._check(self.lib.chimera_engine_set_graph(self._h, int(bool(on))))

    def set_lazy_tail(self, on=True):
        """``step`` leaves the gather + push that closes its last step pending so that the next ``step`` call runs it
        inside its first fused kernel (default on); any other call on the engine completes it first."""
        self._check(self.lib.chimera_engine_set_lazy_tail(self._h, int(bool(on))))

    def graph_info(self):
        """(number of cached step graphs, state: 1 = in use, 0 = not warmed up, -1 = capture failed, graphs off)"""
        n, st = ctypes.c_int(0), ctypes.c_int(0)
        self._check(self.lib.chimera_engine_graph_info(self._h, ctypes.byref(n), ctypes.byref(st)))
        return n.value, st.value

    def use_stream(self, cuda_stream_ptr):
        self._check(self.lib.chimera_engine_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr)))

    def _allreduce_grids(self):
        names = ("J", "Rho") if (self.cfg.space_charge or self.cfg.static_kick) else ("J",)
        for n in names:
            self._comm("all_reduce_" + n, lambda a, n=n: self._dist.all_reduce(self.device_tensor(n), group=self._group))

    def _allreduce_grids_async(self):
        """J then Rho on the collective stream; returns the two handles: fb_in_J only needs J, so the reduction of Rho
        runs behind it"""
        w_j = self._dist.all_reduce(self.device_tensor("J"), group=self._group, async_op=True)
        w_r = None
        if self.cfg.space_charge or self.cfg.static_kick:
            w_r = self._dist.all_reduce(self.device_tensor("Rho"), group=self._group, async_op=True)
        return w_j, w_r

    def _fields_out(self):
        """``Solver.G2B_FBRot`` + ``fb_fld_out`` (solvers.py:536, 450); on kx-slab engines the backward DHT runs
        on this rank's rows, the slabs are all-gathered over NVLink and every rank finishes with the x-FFT.
        Across ranks the E and B halves are pipelined: the all-gather of E runs while B (curl + backward DHT) is
        computed, the all-gather of B while E gets its inverse x-FFT."""
        if not self.slab:
            self.run("fields_out")
            return
        if self.world > 1 and self.colflow:
            c = self.cfg
            self.run("fields_out_a")
            h = self._comm("all_to_all_EB", lambda a: self._all_to_all(self.device_tensor("EB_recv")[:2 * c.nx * self._cb["EB"]],
                                                                      self._cview("EB_slab", c.nx_slab, "EB"), a))
            h.wait()
            self.run("col_bwd")
            h = self._comm("all_gather_EB", lambda a: self._all_gather_into(self._cview("EB", c.nx, "EB"),
                                                                         self.device_tensor("EB_blk")[:2 * c.nx * self._cb["EB"]], a))
            if h is not None:
                h.wait()
            self.run("eb_finish")
            return
        if self.world == 1 or not self.overlap:
            self.run("fields_out_a")
            self._allgather_eb()
            self.run("fields_out_b")
            return
        slab, gath = self._eb_slab(), self.device_tensor("EB_gath")
        hs, hg = slab.numel() // 2, gath.numel() // 2
        self.run("fields_out_a", 1.0)
        w_e = self._all_gather_into(gath[:hg], slab[:hs], async_op=True)
        self.run("fields_out_a", 2.0)
        w_b = self._all_gather_into(gath[hg:], slab[hs:], async_op=True)
        w_e.wait()
        self.run("fields_out_b", 1.0)
        w_b.wait()
        self.run("fields_out_b", 2.0)

    def _all_gather_into(self, out, inp, async_op=False):
        """``all_gather_into_tensor``; on a backend without it (gloo: two ranks sharing one GPU in the tests) the list form"""
        try:
            return self._dist.all_gather_into_tensor(out, inp, group=self._group, async_op=async_op)
        except (RuntimeError, NotImplementedError):
            parts = list(out.view(self.world, -1).unbind(0))
            return self._dist.all_gather(parts, inp.reshape(-1), group=self._group, async_op=async_op)

    def _allgather_eb(self):
        self._all_gather_into(self.device_tensor("EB_gath"), self._eb_slab())

    def _eb_slab(self):
        """the backward-transformed slab (nx_slab, Nr, M, 6) without the spare columns of its allocation"""
        c = self.cfg
        return self.device_tensor("EB_slab")

    def _deposit_only(self):
        self.run("deposit_J")
        if self.cfg.space_charge or self.cfg.static_kick:
            # the background charge enters the sum once (rank 0), chimera_main.py:189-190
            self.run("deposit_rho", 1.0 if self.rank == 0 else 0.0)

    def _deposit_and_reduce(self):
        self._deposit_only()
        if self.world > 1:
            self._allreduce_grids()

    # -- collectives (with the fall-backs the gloo test transport needs) -------------------------------------------
    class _Done:
        def wait(self):
            return True

    def _comm(self, name, fn):
        """run a collective; with ``comm_profile`` on, synchronously between two CUDA events (its own time, no overlap)"""
        if not getattr(self, "comm_profile", False):
            return fn(self.overlap)
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h = fn(False)
        if h is not None:
            h.wait()
        e1.record()
        self._comm_events.append((name, e0, e1))
        return self._Done()

    def comm_timings(self, reset=True):
        """{collective: (ms, calls)} measured while ``comm_profile`` was on"""
        import torch

        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self._comm_events:
            ms, n = out.get(name, (0.0, 0))
            out[name] = (ms + e0.elapsed_time(e1), n + 1)
        if reset:
            self._comm_events = []
        return out

    def _reduce_scatter(self, out, inp, async_op):
        try:
            return self._dist.reduce_scatter_tensor(out, inp, group=self._group, async_op=async_op) or self._Done()
        except (RuntimeError, NotImplementedError):  # gloo: all-reduce, keep the own block
            self._dist.all_reduce(inp, group=self._group)
            out.copy_(inp.view(self.world, -1)[self.rank])
            return self._Done()

    def _all_to_all(self, out, inp, async_op):
        try:
            return self._dist.all_to_all_single(out, inp, group=self._group, async_op=async_op) or self._Done()
        except (RuntimeError, NotImplementedError):  # gloo on CUDA tensors: gather everything, keep what is addressed here
            import torch

            tmp = torch.empty(self.world * inp.numel(), dtype=inp.dtype, device=inp.device)
            self._all_gather_into(tmp, inp)
            out.view(self.world, -1).copy_(tmp.view(self.world, self.world, -1)[:, self.rank])
            return self._Done()

    def _setup_colflow(self):
        """column-block dataflow of the multi-rank solve (csrc/engine.cu): reduce-scatter of J / Rho by
        column block, x-FFT of the own block, all-to-all to kx slabs; backward the other way round and an all-gather of
        the column blocks of EB.  Needs kx slabs over all the ranks of the group."""
        self.colflow = False
        self._comm_events = []
        if not (self.slab and self.world > 1 and self.slab_world == self.world and self.slab_rank == self.rank):
            return
        if not hasattr(self.lib, "chimera_engine_set_colflow") or self.world > 64:
            return
        self._check(self.lib.chimera_engine_set_colflow(self._h, self.rank, self.world))
        c, w = self.cfg, self.world
        self._cb = {k: -(-(c.nrn * c.nm * n) // w) for k, n in (("J", 3), ("Rho", 1), ("EB", 6))}
        self.colflow = True

    def _cview(self, name, rows, key):
        """the first rows x (column block x world) complex entries of a named array, as doubles"""
        return self.device_tensor(name, raw=True)[:2 * rows * self._cb[key] * self.world]

    def _reduce_and_transform(self):
        """sum the deposited grids over the ranks and take them to Fourier-Bessel space on this engine's kx rows
        (``fb_curr_in`` / ``fb_dens_in`` + ``FBGradDens``, solvers.py:407-448)"""
        c = self.cfg
        rho = bool(c.space_charge or c.static_kick)
        if self.world == 1 or not self.colflow:
            w_r = None
            if self.world > 1:
                if self.overlap:
                    w_j, w_r = self._allreduce_grids_async()
                    w_j.wait()
                else:
                    self._allreduce_grids()
            self.run("fb_in_J")
            if w_r is not None:
                w_r.wait()
            if rho:
                self.run("fb_in_rho")
            return
        nx, L, W = c.nx, c.nx_slab, self.world
        blk = lambda name, key: self.device_tensor(name)[:2 * nx * self._cb[key]]  # noqa: E731
        h_j = self._comm("reduce_scatter_J", lambda a: self._reduce_scatter(blk("J_blk", "J"), self._cview("J", nx, "J"), a))
        h_r = self._comm("reduce_scatter_Rho", lambda a: self._reduce_scatter(blk("Rho_blk", "Rho"), self._cview("Rho", nx, "Rho"), a)) if rho else None
        h_j.wait()
        self.run("col_fwd", 0.0)
        a_j = self._comm("all_to_all_J", lambda a: self._all_to_all(self._cview("J_in", L, "J"), blk("J_send", "J"), a))
        a_r = None
        if rho:
            h_r.wait()
            self.run("col_fwd", 1.0)
            a_r = self._comm("all_to_all_Rho", lambda a: self._all_to_all(self._cview("Rho_in", L, "Rho"), blk("Rho_send", "Rho"), a))
        a_j.wait()
        self.run("fb_in_col", 0.0)
        if rho:
            a_r.wait()
            self.run("fb_in_col", 1.0)

    def _static_fields(self):
        """``update_fields`` under 'StaticKick' (chimera_main.py:118-125) across ranks / on kx slabs: the mean momentum of
        every species from its all-reduced moments, then the quasi-static field on this engine's kx rows"""
        px = []
        for sid in range(self.nspecies):
            m = self.beam_moments(sid)
            px.append(m[5] / m[0] if m[0] != 0.0 else float("nan"))
        arr = (ctypes.c_double * len(px))(*px)
        self._check(self.lib.chimera_engine_set_static_px(self._h, arr, len(px)))
        self.run("static_fields")

    def deposit_background(self):
        """``ChimeraRun.dep_bg`` (chimera_main.py:220-248): still species -> BckGrndRho."""
        self.run("deposit_bg")
        if self.world > 1:
            self._dist.all_reduce(self.device_tensor("BckGrndRho"), group=self._group)

    def make_halfstep(self, px0=None, background=False):
        """``ChimeraRun.make_halfstep`` (chimera_main.py:61-80): bin, deposit, static field of the initial
        momenta ``px0`` (one entry per species: ``MomentaMeans[0]``, chimera_main.py:73-75; default: the values given
        to ``add_species``), gather, half Boris push."""
        if px0 is None:
            px0 = tuple(self._px0) or (0.0,)
        self.run("sort", 0.0)
        if background:
            self.deposit_background()
        self._deposit_only()
        self._reduce_and_transform()
        if self.cfg.space_charge or self.cfg.static_kick:
            for p in px0:  # solvers.py:333-358, one static kick per species
                c1, c2 = self.setup.static_coeffs(p)
                self.upload("CPSATD1", c1)
                self.upload("CPSATD2", c2)
                self.run("init_push")
        self._fields_out()
        self.set_time(0.0)  # make_device() (chimera_main.py:78)
        self.run("gather_push", 0.5)

    def step(self, nsteps=1):
        """``nsteps`` x ``ChimeraRun.make_step`` (chimera_main.py:82-92)."""
        s1, s2 = getattr(self, "_win", (0.0, 0.0))
        if s1 or s2:  # a frame that moves every step: keep the Python-side window in step with the C side
            self.cfg.leftX += nsteps * (s1 + s2)
            self.cfg.rightX += nsteps * (s1 + s2)
        if self.world == 1 and not self.slab:
            self._check(self.lib.chimera_engine_step(self._h, _i64(self.istep + 1), _i64(nsteps)))
            self.istep += nsteps
            return
        # same schedule as chimera_engine_step (csrc/engine.cu): inside a multi-step call the particle work between
        # two field solves (gather + push of step k, push_coords + deposits of step k+1) is one fused kernel
        c = self.cfg
        gather_pending = False
        for _ in range(nsteps):
            self.istep += 1
            sort_now = c.sort_every > 0 and self.istep % c.sort_every == 0
            self.set_time((self.istep - 1) * c.dt)  # the pending gather + push closes the previous step
            bg = 1.0 if self.rank == 0 else 0.0  # the background charge enters the all-reduced density once
            if gather_pending and not sort_now and self.fuse and not c.static_kick:
                self.run("particles_fused", bg)  # ('StaticKick' deposits rho on coords_halfstep: not the fused kernel's layout)
            elif self.fuse:  # re-binning step / first step of the call: one kernel before the sort, one after
                if gather_pending:
                    self.run("gather_push_coords")
                else:
                    self.run("window", 1.0)
                    self.run("push_coords")
                if sort_now:
                    self.run("sort", 1.0)
                self.run("deposit_fused", bg)
            else:
                if gather_pending:
                    self.run("gather_push", 1.0)
                win = any(getattr(self, "_win", (0.0, 0.0)))
                if win:
                    self.run("window", 1.0)
                self.run("push_coords")
                if sort_now:
                    self.run("sort", 1.0)
                self.run("deposit_J")
                if win:
                    self.run("window", 2.0)
                if c.space_charge or c.static_kick:
                    self.run("deposit_rho", bg)
            self._reduce_and_transform()
            if c.static_kick:
                self._static_fields()
            else:
                self.run("poisson")
                self.run("maxwell")
            self._fields_out()
            gather_pending = True
        if gather_pending:
            self.set_time(self.istep * c.dt)
            self.run("gather_push", 1.0)

    # -- host-buffer stepping ------------------------------------------------------------------
    def step_host(self, coords, coords_half, momenta, weights, EG_fb=None, gradRho_fb_nxt=None, sid=0, rebin=False):
        """One ``ChimeraRun.make_step`` (chimera_main.py:82-92) on HOST arrays, the reference's calling model.

        ``coords``/``momenta`` (3,Np) and ``weights`` (Np,) are the species' numpy arrays (Fortran order,
        float64), updated in place; ``coords_half`` (3,Np) receives the centred positions (``None``: not copied back --
        they are only used inside the step, for the deposit and the re-binning the engine does itself); ``EG_fb`` and
        ``gradRho_fb_nxt`` are the solver's spectral state, updated in place (``None``: keep the
        engine-resident copy).  Host<->device copies are pipelined with the kernels inside the call
        (csrc/engine.cu ``chimera_engine_step_host``); page-lock the arrays once with :meth:`pin` for full
        PCIe speed.  Returns the number of particles kept (the first entries of the arrays are valid)."""
        n = coords.shape[1]
        for name, arr, shp in (("coords", coords, (3, n)), ("coords_half", coords_half, (3, n)),
                               ("momenta", momenta, (3, n)), ("weights", weights, (n,))):
            if arr is None and name == "coords_half":
                continue
            if arr.dtype != np.float64 or not arr.flags.f_contiguous or arr.shape != shp or not arr.flags.writeable:
                raise ValueError("step_host: %s must be a writeable Fortran-ordered float64 array of shape %r" % (name, shp))
        for name, arr in (("EG_fb", EG_fb), ("gradRho_fb_nxt", gradRho_fb_nxt)):
            if arr is not None and (arr.dtype != np.complex128 or not arr.flags.f_contiguous or arr.shape != self.shape_of(name)):
                raise ValueError("step_host: %s must be a Fortran-ordered complex128 array of shape %r" % (name, self.shape_of(name)))
        self.istep += 1
        n_out = _i64(0)
        vp = lambda a: ctypes.c_void_p(a.ctypes.data if a is not None else None)  # noqa: E731
        if self.world == 1:
            self._check(self.lib.chimera_engine_step_host(
                self._h, int(sid), vp(coords), vp(coords_half), vp(momenta), vp(weights), _i64(n), ctypes.byref(n_out),
                vp(EG_fb), vp(gradRho_fb_nxt), _i64(self.istep), int(bool(rebin))))
            return n_out.value
        # particles sharded over the ranks: deposit locally, sum the grids over NVLink, then the field update
        self._check(self.lib.chimera_engine_set_rho_from_bg(self._h, int(self.rank == 0)))
        self._check(self.lib.chimera_engine_step_host_begin(
            self._h, int(sid), vp(coords), vp(coords_half), vp(momenta), vp(weights), _i64(n), vp(EG_fb),
            vp(gradRho_fb_nxt), _i64(self.istep), int(bool(rebin))))
        self._allreduce_grids()
        if self.slab:
            self._check(self.lib.chimera_engine_step_host_mid(self._h))
            self._allgather_eb()
        self._check(self.lib.chimera_engine_step_host_end(self._h, ctypes.byref(n_out)))
        return n_out.value

    def pin(self, *arrays):
        """Page-lock numpy arrays (cudaHostRegister) so that step_host's copies run at PCIe speed."""
        for a in arrays:
            if a is not None and a.nbytes:
                self._check(self.lib.chimera_host_register(ctypes.c_void_p(a.ctypes.data), _i64(a.nbytes)))
                self._pinned.append(a)

    def unpin_all(self):
        for a in self._pinned:
            self.lib.chimera_host_unregister(ctypes.c_void_p(a.ctypes.data))
        self._pinned = []

    # -- profiling ---------------------------------------------------------------------------
    def profile(self, on=True):
        self._check(self.lib.chimera_engine_profile(self._h, int(on)))

    def timings(self, reset=True):
        ms = (ctypes.c_double * len(PHASES))()
        calls = (_i64 * len(PHASES))()
        self._check(self.lib.chimera_engine_timings(self._h, ms, calls, int(reset)))
        return {n: (ms[i], calls[i]) for i, n in enumerate(PHASES) if calls[i]}
