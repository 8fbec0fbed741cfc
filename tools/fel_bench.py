"""BASELINE.json configs[3], one GPU's share: unaveraged FEL undulator beam, envelope solver (KxShift), one
azimuthal mode, NoPoissonCorrection, analytic undulator, 1.25e8 macro-particles (1e9 over 8 GPUs).  Device-
resident engine with the 'Staged' window moving every step inside it, CUDA events around K make_steps; per-phase
times from the engine profile.  The beam fills
x in +-lbx/2, r < lbr of the fel-testrun geometry (doc/tests/fel-testrun.py:12-58) on an Nx x Nr = 2048 x 256
grid: ~7600 particles per occupied cell, i.e. the deposit is dominated by same-cell accumulation.
Writes gpurun_out/fel_bench.json.

  python tools/fel_bench.py [--np 1.25e8] [--steps 10] [--warmup 3]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.path.insert(0, ROOT)
from chimera_b200.engine import Engine  # noqa: E402
from chimera_b200.solver_setup import SolverSetup  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--np", type=float, default=1.25e8)
ap.add_argument("--nx", type=int, default=2048)
ap.add_argument("--nr", type=int, default=256)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--sort-every", type=int, default=None,
                help="re-binning cadence in steps (default: the reference's Xchunked[1]+1 = 7, chimera_main.py:310); the "
                     "co-moving beam keeps its cell order much longer, so production runs can re-bin rarely")
ap.add_argument("--fused-profile", action="store_true", help="per-stage clock shares of the fused particle kernel (diagnosis)")
a = ap.parse_args()

K0, lam0, periods = 1.95, 2.8, 10
g0 = 200 / 0.511
gg = g0 / (1.0 + K0 ** 2 / 2) ** 0.5
k_res, vb = 2 * gg ** 2, (1.0 - gg ** -2) ** 0.5
Lgx, Rg, Rcut = 200e-4 / lam0, 1000e-4 / lam0, 700e-4 / lam0
lbx = lbr = 80e-4 / lam0
dt = 1.0 / 30
cfg = {"Grid": (-0.5 * Lgx, 0.5 * Lgx, Rg, Lgx / a.nx, Rg / a.nr), "TimeStep": dt, "MaxAzimuthMode": 0,
       "KxShift": k_res, "Rcut": Rcut, "CoPropagative": vb, "Xchunked": (16, 6),
       "Features": {"NoPoissonCorrection": True}}
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
group = None
if world > 1:  # one rank per GPU: particles sharded (weak scaling), J all-reduced, spectral solve by kx slab
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = True
from chimera_b200 import _lib  # noqa: E402

_lib.load().chimera_set_device(local)
S = SolverSetup(cfg)
eng = Engine(S, group=group, sort_every=a.sort_every)
eng.use_stream(torch.cuda.current_stream().cuda_stream)
eng.add_device("undul_analytic", np.array([K0, 1.0, 1.0, float(periods)]))
# MovingFrame {'TimeStep': dt, 'Steps': 1, 'Velocity': vb, 'Features': ('Staged', 'NoSorting')} (fel-testrun.py:61-63):
# the window follows the beam, half a shift before push_coords and half between dep_curr and dep_dens
eng.set_window(vb, time_step=dt, staged=True)

n = int(a.np)
g = torch.Generator(device="cuda")
g.manual_seed(20260101 + rank)
rnd = lambda *s: torch.rand(*s, device="cuda", dtype=torch.float64, generator=g)  # noqa: E731
rndn = lambda *s: torch.randn(*s, device="cuda", dtype=torch.float64, generator=g)  # noqa: E731
x = (rnd(n) - 0.5) * lbx
r = lbr * torch.sqrt(rnd(n))  # uniform in the disc
th = 2 * np.pi * rnd(n)
coords = torch.stack((x, r * torch.sin(th), r * torch.cos(th)), dim=1).contiguous()
mom = torch.stack((g0 + 1e-4 * g0 * rndn(n), 2e-5 * g0 * rndn(n), 2e-5 * g0 * rndn(n)), dim=1).contiguous()
dens = 20e-12 / 1.6022e-19 / (np.pi * 80e-4 ** 3) / (1.1e21 / 2.8e4 ** 2)
w = torch.full((n,), -dens * np.pi * lbr ** 2 * lbx / n, device="cuda", dtype=torch.float64)
torch.cuda.synchronize()
eng.add_species_device(coords.data_ptr(), mom.data_ptr(), w.data_ptr(), n)
del x, r, th, coords, mom, w
torch.cuda.empty_cache()
eng.make_halfstep(px0=(0.0,))
eng.step(a.warmup)
eng.sync()
eng.profile(True)
eng.timings(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


barrier()
e0.record()
eng.step(a.steps)
e1.record()
barrier()
ms = e0.elapsed_time(e1) / a.steps
if world > 1:
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
ph = eng.timings(reset=True)
eng.profile(False)
kept = eng.count(0)
if world > 1:
    c = torch.tensor([float(kept)], device="cuda", dtype=torch.float64)
    dist.all_reduce(c)
    kept = int(c.item())
hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
per = {k: v[0] / v[1] for k, v in ph.items() if v[1]}
out = {"workload": "FEL undulator beam, envelope solver Nx=%d Nr=%d 1 mode, undul_analytic K0=1.95, %.3g macro-particles, "
                   "Xchunked=(16,6), NoPoissonCorrection" % (a.nx, a.nr, n),
       "metric": "particle-steps/s full PIC cycle", "value": kept / (ms * 1e-3), "ms_per_step": ms, "steps": a.steps,
       "warmup": a.warmup, "sort_every": eng.cfg.sort_every, "n_gpus": world, "particles_per_gpu": n, "particles_after": kept, "phases_ms_per_call": per,
       "phase_calls": {k: v[1] for k, v in ph.items()},
       # envelope cycle, SURVEY 8d basis: push_coords 96 + dep_curr_env 56 + proj_fld_env 128 + push_velocs 96 B
       "particle_cycle_alg_bytes_per_gpu": 376.0 * kept / world, "hbm_gbs_peak": hbm}
if "particles_fused" in per:
    out["fused_frac_of_hbm"] = 376.0 * kept / world / (per["particles_fused"] * 1e-3) / 1e9 / hbm
if a.fused_profile:
    import ctypes

    lib = eng.lib
    lib.chimera_fused_profile(1)
    eng.step(4)
    eng.sync()
    cyc = (ctypes.c_ulonglong * 8)()
    lib.chimera_fused_profile_read(cyc)
    lib.chimera_fused_profile(0)
    tot = float(sum(cyc[:6])) or 1.0
    out["fused_stages"] = {nm: cyc[i] / tot for i, nm in enumerate(
        ("A_records_histogram", "BC_scan_sort", "D_gather", "E_push", "F_deposit", "G_cell_changers"))}
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = "" if a.sort_every is None else "_sort%d" % a.sort_every
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fel_bench_n%d%s.json" % (world, tag)), "w"), indent=1)
    print(json.dumps(out))
eng.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
