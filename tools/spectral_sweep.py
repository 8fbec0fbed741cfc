"""BASELINE.json configs[4]: DHT / PSATD microbenchmark sweep, Nr 256-2048, Nkz (= Nx) 1024-16384, 1-3 azimuthal
modes, against the FP64-tensor and HBM rooflines.

For each shape a device-resident engine without particles runs the spectral half of make_step
(fb_in_J, fb_in_rho + fb_grad, 3 x (fb_graddiv + poiss_corr), maxwell_push_with_spchrg, fb_rot + fb_eb_out +
eb_correction) on random spectral data; CUDA events per phase (engine profile) and per contraction launch (gemm
profile).  Reported per shape: ms per step of the whole spectral update, TFLOP/s of the DMMA contraction kernel
(4 Nx K N flop per contraction) and its fraction of the cuBLAS DGEMM peak measured in the same
process.  Writes gpurun_out/spectral_sweep.json.

  python tools/spectral_sweep.py [--quick]"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chimera_b200 import _lib, synthetic  # noqa: E402
from chimera_b200.engine import Engine  # noqa: E402
from chimera_b200.solver_setup import SolverSetup  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()


def dgemm_peak(n=8192):
    x = torch.randn(n, n, dtype=torch.float64, device="cuda")
    y = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        x @ y
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); x @ y; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * n ** 3 / (best * 1e-3) / 1e12


lib = _lib.load()
peak = dgemm_peak()
hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
shapes = [(1024, 256, 1), (4096, 512, 3)] if a.quick else [
    (nx, nr, m) for nr in (256, 512, 1024, 2048) for nx in (1024, 4096, 16384) for m in (1, 3)
    # one (Nx, Nkr, M, 6) complex array is 96 Nx Nr M bytes; the engine holds ~12 of that size
    if 96 * nx * nr * m * 14 < 120e9]
out = {"gpu": torch.cuda.get_device_name(0), "cublas_dgemm_tflops": peak, "hbm_gbs": hbm, "cases": []}
rng = np.random.default_rng(1)
for nx, nr, m in shapes:
    t0 = time.time()
    S = SolverSetup(synthetic.lwfa_solver_config(nx=nx, nr=nr, modes=m))
    eng = Engine(S)
    eng.use_stream(torch.cuda.current_stream().cuda_stream)
    for name in ("EG_fb", "J", "Rho"):
        t = eng.device_tensor(name)
        t.copy_(torch.randn(t.shape, dtype=torch.float64, device="cuda"))
    for name in ("gradRho_fb_nxt", "gradRho_fb_prv"):
        t = eng.device_tensor(name)
        t.copy_(torch.randn(t.shape, dtype=torch.float64, device="cuda") * 1e-3)

    def spectral_step():
        eng.run("fb_in_J"); eng.run("fb_in_rho"); eng.run("poisson"); eng.run("maxwell"); eng.run("fields_out")

    spectral_step()
    eng.sync()
    eng.profile(True); eng.timings(reset=True)
    lib.chimera_gemm_profile(1)
    lib.chimera_gemm_profile_read(None, None, None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        # keep the fields bounded: the random J would pump EG_fb up over the iterations
        spectral_step()
    e1.record(); torch.cuda.synchronize()
    eng.sync()
    ms, fl, nl = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
    lib.chimera_gemm_profile_read(ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(nl), 1)
    lib.chimera_gemm_profile(0)
    ph = eng.timings(reset=True)
    eng.profile(False)
    per = {k: v[0] / max(v[1], 1) for k, v in ph.items() if v[1]}
    step_ms = sum(per.get(k, 0.0) for k in ("fb_in_J", "fb_in_rho", "poisson", "maxwell", "fields_out"))
    tf = fl.value / (ms.value * 1e-3) / 1e12 if ms.value > 0 else 0.0
    pts = float(np.prod(S.shape_fb))  # spectral points (Nx, Nkr, M)
    case = {"Nx": nx, "Nr": nr, "modes": m, "dht_psatd_ms_per_step": step_ms, "phases_ms": per,
            "gemm_ms_per_step": ms.value / a.iters, "gemm_launches_per_step": nl.value / a.iters,
            "gemm_tflops": tf, "gemm_frac_of_dgemm_peak": tf / peak,
            "maxwell_push_gbs": 416.0 * pts / (per.get("maxwell", 1e9) * 1e-3) / 1e9,
            "maxwell_push_frac_hbm": 416.0 * pts / (per.get("maxwell", 1e9) * 1e-3) / 1e9 / hbm,
            "wall_ms_per_step_with_profile_events": e0.elapsed_time(e1) / a.iters, "setup_s": time.time() - t0}
    out["cases"].append(case)
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in case.items() if k != "phases_ms"}, flush=True)
    eng.close()
    del eng
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "spectral_sweep.json"), "w"), indent=1)
