#!/usr/bin/env python
"""Benchmark of the PIC-cycle hot path (BASELINE.json metric: particle-steps/s, full PIC cycle).

  python bench.py [--gpus N] [--steps K] [--warmup W]                 our CUDA engine
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] the reference's algorithm on the
                                                                       host cores (oracle port: the
                                                                       Fortran cannot be built here)
Under torchrun (N > 1) one rank per GPU: particles are sharded (weak scaling: the per-GPU particle
count is fixed), the deposited grids are all-reduced over NCCL.  One JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...", printed when NCCL_DEBUG is set) goes
# to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

METRIC = "particle-steps/s full PIC cycle"
UNIT = "particle-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c1a", "c1b", "c2_static", "c2_pic"],
                    help="c3 = BASELINE configs[2], the LWFA synthetic case the metric is quoted on (default); c1a / c1b = the "
                         "FEL and LPA stages of configs[0] (the shipped demo), c2_* = the two stages of configs[1] (space-charge "
                         "drift) at their own small grids, one GPU")
    ap.add_argument("--ppc", type=int, default=48, choices=[16, 48],
                    help="macro-particles per cell: 48 = 1.0e8 particles (the '~1e8' BASELINE.json names), 16 = 3.4e7")
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--nr", type=int, default=512)
    ap.add_argument("--modes", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-per-call", action="store_true", help="skip the per-function drop-in timing (e2e_per_call)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fused-profile", action="store_true",
                    help="after the timed run: per-stage clock shares of the fused particle kernel (adds barriers; diagnosis)")
    return ap.parse_args()


CELL = {16: (2, 2, 4), 48: (4, 3, 4)}


def workload_name(a):
    return ("LWFA synthetic Nz=%d Nr=%d %d azimuthal modes, %d ppc (%.2e macro-particles per GPU), real PSATD solver "
            "with SpaceCharge + 3 Poisson iterations, Xchunked=(16,10), dt=dx" % (a.nx, a.nr, a.modes, a.ppc, n_particles(a)))


def n_particles(a):
    return (a.nx - 1 - 24) * (a.nr - 8) * a.ppc


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        time.sleep(0.15)
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_substr):
    """DRAM bytes per launch of a kernel from the committed ncu --set full capture (profiles/traffic.json,
    written by tools/ncu_summary.py traffic); None when the kernel is not in the capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    for name, v in json.load(open(p)).get("kernels", {}).items():
        if kernel_substr in name:
            return v["dram_bytes_per_launch"]
    return None


def dgemm_peak_tflops(torch, n=8192, reps=5):
    """cuBLAS DGEMM n^3: the FP64-tensor denominator (MEASURED_PEAKS.json has no FP64 figure)."""
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


# algorithmic HBM bytes per particle of the particle phases of the engine
ALG_BYTES = {
    "push_coords": 96.0,   # R x,p  W x,x_half
    "deposit_J": 56.0,     # R x_half,p,w   (+ grid RW, added below)
    "deposit_rho": 32.0,   # R x,w
    "gather_push": 80.0,   # fused proj_fld + push_velocs: R x,w,p  W p  (the per-particle EB never hits HBM)
    # the fused kernel does the whole particle cycle of one step (push_coords + dep_curr + dep_dens + proj_fld +
    # push_velocs): SURVEY.md section 8d counts that as 408 B/particle-step over the reference-API arrays
    # (96 + 56 + 32 + 128 + 96); fused and resident it only has to move 128 B (R x,p,w  W x,x_half,p)
    "particles_fused": 408.0,
    # the two halves of a re-binning step: gather + push + push_coords before the sort (proj_fld 128 + push_velocs 96 +
    # push_coords 96), dep_curr + dep_dens after it (56 + 32)
    "gather_push_coords": 320.0,
    "deposit_fused": 88.0,
}
FUSED_RESIDENT_BYTES = 128.0


TRAFFIC_KERNEL = {"particles_fused": "fused_particles_k", "gather_push_coords": "gather_push_coords_k", "gather_push": "gather_push_binned_k",
                  "deposit_J": "deposit_binned_k<0, 1", "deposit_rho": "deposit_binned_k<0, 0", "push_coords": "push_coords_k"}


# ------------------------------------------------------------------------------------------------
def build_problem(a, torch, rank, world, group):
    from chimera_b200 import synthetic
    from chimera_b200.engine import Engine
    from chimera_b200.solver_setup import SolverSetup
    import chimera_b200.fimera as gfim

    S = SolverSetup(synthetic.lwfa_solver_config(nx=a.nx, nr=a.nr, modes=a.modes))
    eng = Engine(S, group=group)
    eng.use_stream(torch.cuda.current_stream().cuda_stream)
    x, p, w = synthetic.plasma_fixed_cell(S.Args, cell=CELL[a.ppc], seed=20260101 + rank, xp=torch)
    n = x.shape[0]
    torch.cuda.synchronize()
    eng.add_species_device(x.data_ptr(), p.data_ptr(), w.data_ptr(), n)
    del x, p, w
    torch.cuda.empty_cache()
    # still ions coincide with the electrons at t=0: BckGrndRho = - rho_e(0)  (chimera_main.py:220 dep_bg)
    eng.run("sort", 0.0)
    eng.run("deposit_rho", 0.0)
    if world > 1:
        torch.distributed.all_reduce(eng.device_tensor("Rho"), group=eng._group)
    bg = eng.device_tensor("BckGrndRho")
    bg.copy_(-eng.device_tensor("Rho"))
    eng.upload("EG_fb", synthetic.laser_seed(S, gfim))
    eng.make_halfstep(px0=(0.0,))
    eng.sync()
    return S, eng, n


def bind_to_gpu_numa_node(local):
    """Pin this rank to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function) so that the
    host arrays of the end-to-end leg are first-touched on the GPU's NUMA node.  Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # 00000000:1b:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        cpus = set()
        for part in open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_ours(a):
    import torch

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        bind_to_gpu_numa_node(local)
    group = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = True
    from chimera_b200 import _lib

    lib = _lib.load()
    lib.chimera_set_device(local)
    S, eng, n_local = build_problem(a, torch, rank, world, group)

    def barrier():
        eng.sync()  # completes the gather + push a step call leaves pending (chimera_engine_set_lazy_tail)
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    eng.step(a.warmup)
    barrier()
    eng.profile(True)
    eng.timings(reset=True)
    lib.chimera_gemm_profile(1)
    l0 = _lib.kernel_launches()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    eng.step(a.steps)
    eng.sync()  # K complete steps inside the timed region: the last step's gather + push included
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.finish() if sampler else None
    launches = _lib.kernel_launches() - l0
    phases = eng.timings(reset=True)
    eng.profile(False)
    g_ms, g_fl, g_n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
    lib.chimera_gemm_profile_read(ctypes.byref(g_ms), ctypes.byref(g_fl), ctypes.byref(g_n), 1)
    lib.chimera_gemm_profile(0)
    n_now = eng.count(0)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_total = float(t.item())
        cnt = torch.tensor([float(n_local)], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(cnt)
        n_total = int(cnt.item())
    else:
        n_total = n_local
    ms_step = ms_total / a.steps
    value = n_total / (ms_step * 1e-3)

    out = None
    if rank == 0:
        hbm_peak, hbm_src = measured_peaks()
        fp64_peak = dgemm_peak_tflops(torch)
        c = eng.cfg
        grid_pts = c.nx * c.nrn * c.nm
        stages = {}
        for name, (ms, calls) in phases.items():
            per = ms / calls
            ent = {"ms_per_call": per, "share": ms / ms_total}
            if name in ALG_BYTES:
                by = ALG_BYTES[name] * n_local
                if name == "particles_fused":
                    by += (96.0 + 2 * 48.0 + 2 * 16.0) * grid_pts  # EB read once, J and Rho read-modify-write
                if name == "deposit_fused":
                    by += (2 * 48.0 + 2 * 16.0) * grid_pts
                if name == "gather_push_coords":
                    by += 96.0 * grid_pts
                if name == "deposit_J":
                    by += 2 * 48.0 * grid_pts
                if name == "deposit_rho":
                    by += 2 * 16.0 * grid_pts
                ent.update(bound="hbm", alg_bytes=by, achieved_gbs=by / (per * 1e-3) / 1e9,
                           frac=by / (per * 1e-3) / 1e9 / hbm_peak)
                if name == "particles_fused":  # the stricter figure: bytes a fused, resident kernel cannot avoid
                    by2 = by - (ALG_BYTES[name] - FUSED_RESIDENT_BYTES) * n_local
                    ent.update(fused_resident_bytes=by2, frac_fused_resident=by2 / (per * 1e-3) / 1e9 / hbm_peak)
            stages[name] = ent
        gemm = {"ms_per_launch": g_ms.value / max(g_n.value, 1), "launches_per_step": g_n.value / a.steps,
                "share": g_ms.value / ms_total, "tflops": g_fl.value / (g_ms.value * 1e-3) / 1e12 if g_ms.value else 0.0}
        # dominant kernel by share of the timed region: the DMMA contraction or one particle kernel
        cand = {k: v["share"] for k, v in stages.items() if "bound" in v}
        top = max(cand, key=cand.get) if cand else None
        if top is None or gemm["share"] >= cand[top]:
            roof = {"kernel": "gemm_dmma_k (DHT + mode-coupling contractions)", "bound": "tensor",
                    "achieved": gemm["tflops"], "peak": fp64_peak, "unit": "TFLOP/s", "frac": gemm["tflops"] / fp64_peak,
                    "traffic": ncu_traffic("gemm_dmma_k"), "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (FP64 tensor; MEASURED_PEAKS.json has no FP64 figure)",
                    "frac_of_nominal_40_tflops": gemm["tflops"] / 40.0,
                    "share_of_step": gemm["share"]}
        else:
            s = stages[top]
            roof = {"kernel": top, "bound": "hbm", "achieved": s["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": s["frac"], "traffic": ncu_traffic(TRAFFIC_KERNEL.get(top, top)), "peak_source": hbm_src,
                    "share_of_step": s["share"]}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "particles_total": n_total, "particles_after": n_now,
                       "parallelism": ("particles sharded x%d; spectral solve sharded by kx slab (%d rows per GPU); %s (NCCL)"
                                       % (world, eng.cfg.nx_slab, "J/Rho reduce-scattered by column block, x-FFT per block, all-to-all "
                                          "to the slabs and back, EB column blocks all-gathered" if eng.colflow else
                                          "J/Rho all-reduced, EB slabs all-gathered")) if eng.slab else
                                      ("particles sharded x%d, grids all-reduced (NCCL), spectral solve replicated" % world),
                       "l2": "inputs larger than L2 (particle arrays %.1f GB, grids %.2f GB)"
                             % (n_local * 80 / 1e9, grid_pts * 16 * 10 / 1e9)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "stages": stages, "gemm": gemm,
            # second half of BASELINE.json's metric: "DHT+PSATD ms/step" = every spectral phase of one step
            "dht_psatd_ms_per_step": sum(v["ms_per_call"] for k, v in stages.items()
                                         if k in ("fb_in_J", "fb_in_rho", "poisson", "maxwell", "fields_out", "fields_out_a", "fields_out_b")),
            "fp64_peak_tflops": fp64_peak,
        }
    if world > 1:
        # the collectives on their own: a short extra pass in which each one runs synchronously between two events
        # (in the timed region above they overlap with the kernels; these are their stand-alone times)
        eng.comm_profile = True
        eng.comm_timings(reset=True)
        eng.step(3)
        comm = eng.comm_timings(reset=True)
        eng.comm_profile = False
        if rank == 0:
            out["collectives"] = {k: {"ms_per_call": ms / n, "calls_per_step": n / 3.0} for k, (ms, n) in comm.items()}
            out["collectives"]["dataflow"] = ("column blocks: reduce-scatter J/Rho -> x-FFT of the own block -> all-to-all to kx slabs; "
                                              "slab -> all-to-all -> inverse x-FFT of the own block -> all-gather EB") if eng.colflow else \
                                             "all-reduce J/Rho, x-FFT of everything on every rank, all-gather of the EB slabs"
    if a.fused_profile and rank == 0:
        lib.chimera_fused_profile(1)
        eng.profile(True)
        eng.timings(reset=True)
        eng.step(3)
        ph = eng.timings(reset=True)
        eng.profile(False)
        cyc = (ctypes.c_ulonglong * 8)()
        lib.chimera_fused_profile_read(cyc)
        lib.chimera_fused_profile(0)
        tot = float(sum(cyc[:6])) or 1.0
        out["fused_stages"] = {n: cyc[i] / tot for i, n in enumerate(
            ("A_records_histogram", "BC_scan_sort", "D_gather", "E_push", "F_deposit", "G_cell_changers"))}
        out["fused_stages"]["cycles_per_cta"] = tot / max(int(cyc[7]), 1)
        out["fused_stages"]["ms_per_call_with_profile_barriers"] = ph["particles_fused"][0] / ph["particles_fused"][1]
    # ---- end to end through the reference-facing drop-in (host buffers, copies inside the timed region)
    if not a.no_e2e:
        e2e, state = run_e2e(a, torch, S, eng, n_local, world)
        if rank == 0:
            out["e2e"] = e2e
            if world == 1 and not a.no_per_call:
                out["e2e_per_call"] = run_e2e_per_call(a, torch, S, eng, state, resident=True)
                out["e2e_per_call_copy"] = run_e2e_per_call(a, torch, S, eng, state, resident=False)
    if rank == 0 and not a.no_cpu and world == 1:
        if a.no_e2e:
            state = eng.particles(0) + (eng.download("EG_fb"), eng.download("gradRho_fb_nxt"))
        out["cpu_baseline"] = cpu_baseline(a, S, state, eng.download("BckGrndRho"))
    eng.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


# ------------------------------------------------------------------------------------------------
def run_e2e(a, torch, S, eng, n_local, world):
    """End to end with HOST buffers: the whole PIC state (particle arrays + the solver's spectral state)
    lives in page-locked host memory, as numpy owns it in the reference; every step goes through the C ABI
    ``chimera_engine_step_host`` which copies the state in, runs make_step and copies the new state out,
    all inside the timed region (copies pipelined with the kernels on separate streams)."""
    from chimera_b200 import _lib

    lib = _lib.load()
    x, xh, p, w = eng.particles(0)
    eg = eng.download("EG_fb")
    g = eng.download("gradRho_fb_nxt") if eng.cfg.space_charge else None
    eng.pin(x, xh, p, w, eg, g)
    n = x.shape[1]
    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # coords_halfstep is not requested back: it is recomputed by every step and only used inside it (deposit,
    # re-binning), which the engine does itself
    n = eng.step_host(x[:, :n], None, p[:, :n], w[:n], eg, g)  # warm-up
    barrier()
    lib.chimera_host_traffic(None, None, 1)
    t = time.perf_counter()
    done = 0
    for _ in range(a.e2e_steps):
        done += n
        n = eng.step_host(x[:, :n], None, p[:, :n], w[:n], eg, g)
    barrier()
    dt = (time.perf_counter() - t) / a.e2e_steps
    h2d, d2h = ctypes.c_longlong(), ctypes.c_longlong()
    lib.chimera_host_traffic(ctypes.byref(h2d), ctypes.byref(d2h), 1)  # counted by the library per copied buffer
    if world > 1:  # max time over the ranks, particles summed
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        cc = torch.tensor([float(done)], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(cc)
        dt, done = float(tt.item()), float(cc.item())
    out = {"value": done / a.e2e_steps / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": a.e2e_steps,
           "h2d_bytes_per_step": int(h2d.value // a.e2e_steps) * world, "d2h_bytes_per_step": int(d2h.value // a.e2e_steps) * world,
           "path": "C ABI chimera_engine_step_host (Engine.step_host): coords, momenta, weights, EG_fb, gradRho_fb_nxt "
                   "host->device and coords, momenta, EG_fb, gradRho_fb_nxt device->host every step (coords_halfstep, "
                   "recomputed by every step and used only inside it, is not requested back), page-locked host arrays"}
    eng.unpin_all()
    xh[:, :n] = eng.particles(0)[1][:, :n]  # untimed: the arms that start from this state take every array
    return out, (x[:, :n], xh[:, :n], p[:, :n], w[:n], eg, g)


def run_e2e_per_call(a, torch, S, eng, state, resident):
    """The same step driven call by call through chimera_b200.fimera -- the f2py-compatible drop-in -- in the
    reference's make_step sequence (tests/pic_ref.RefRun == chimera_main.py:82-92, numpy-side statements included):
    ~21 synchronous calls per step.  resident=True: the drop-in's resident mode (chimera_b200/resident.py), the driver's
    numpy arrays in CUDA managed memory and no per-call copies; False: every call copies its arguments in and out."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import copy

    import chimera_b200.fimera as gfim
    from pic_ref import RefRun, RefSpecies
    from chimera_b200 import _lib

    lib = _lib.load()
    x, xh, p, w, eg, g = state
    bck = eng.download("BckGrndRho")
    if resident:
        # what a driver that imports the drop-in with CHIMERA_B200_RESIDENT=1 gets: every array it creates afterwards
        # (solver tables, grids, particle arrays) comes from the managed allocator
        gfim.resident(True)
        S = copy.deepcopy(S)
        x, xh, p, w, eg, g, bck = (np.array(v, order="F") for v in (x, xh, p, w, eg, g, bck))
    sp = RefSpecies.__new__(RefSpecies)
    sp.coords, sp.momenta, sp.weights, sp.coords_halfstep = x, p, w, xh
    sp.push_fact, sp.still, sp.devices, sp.chunks = -2 * np.pi, False, [], eng.chunks(0)
    sp.EB = np.zeros((6, 0), order="F")
    timer = StageTimer(gfim)
    run = RefRun(timer, S, [sp], sort_every=0)
    run.Bck = bck
    run.EG_fb = eg
    run.g_nxt = g
    try:
        for _ in range(2 if resident else 1):  # warm-up (scratch growth, cuFFT plans, first touch of the managed arrays)
            run.make_step()
        torch.cuda.synchronize()
        lib.chimera_host_traffic(None, None, 1)
        timer.t.clear()
        nt = 3 if resident else 1
        t = time.perf_counter()
        for _ in range(nt):
            run.make_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / nt
        h2d, d2h = ctypes.c_longlong(), ctypes.c_longlong()
        lib.chimera_host_traffic(ctypes.byref(h2d), ctypes.byref(d2h), 1)
    finally:
        if resident:
            gfim.resident(False)
    calls = {k: 1e3 * v / nt for k, v in sorted(timer.t.items(), key=lambda kv: -kv[1])}
    calls["python_statements"] = dt * 1e3 - sum(calls.values())
    return {"value": sp.coords.shape[1] / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": nt,
            "h2d_bytes_per_step": int(h2d.value // nt), "d2h_bytes_per_step": int(d2h.value // nt), "calls_ms": calls,
            "path": ("chimera_b200.fimera per-function drop-in, RESIDENT mode: numpy arrays in CUDA managed memory (numpy data "
                     "allocator), no staging copies (the byte counts are what the library still copied: small tables), the "
                     "driver's whole-array statements on the device; every call synchronous") if resident else
                    "chimera_b200.fimera per-function drop-in (pageable numpy buffers, every call copies in and out, synchronous)"}


# ------------------------------------------------------------------------------------------------
# BASELINE configs[0] / configs[1]: the shipped demos at their own (small) grids, one GPU
SMALL_NAMES = {
    "c1a": "fel-lpa-demo FEL stage (doc/tests/fel-testrun.py): envelope solver Nx=120 Nr=85, 1 mode, analytic undulator, "
           "'Staged' frame every step",
    "c1b": "fel-lpa-demo LPA stage (doc/tests/lpa-testrun.py): real solver Nx=528 Nr=65, 2 modes, SpaceCharge + still ions, a0=3 pulse",
    "c2_static": "space-charge drift demo, 'StaticKick' stage: Nx=304 Nr=301, 2 modes, Gaussian beam px=50, 'Staged' frame every step",
    "c2_pic": "space-charge drift demo, 'SpaceCharge' stage: Nx=304 Nr=301, 2 modes, Gaussian beam px=50, 'Staged' frame every step",
}


def small_build(a, fim, engine):
    """(S, species, eg0, case) and, with engine=True, the resident engine after make_halfstep"""
    import copy

    from chimera_b200 import synthetic
    from chimera_b200.solver_setup import SolverSetup

    c = synthetic.baseline_case(a.config)
    S = SolverSetup(copy.deepcopy(c["cfg"]))
    species = synthetic.baseline_species(S, c)
    eg0 = S.add_gauss_beam(fim, c["laser"]) if c["laser"] else S.zeros_fb(6)
    eng = None
    if engine:
        from chimera_b200.engine import Engine

        dev = c["device"]
        eng = Engine(S, undulator=dict(zip(("a0", "lambda", "X0", "Lx"), dev[1])) if dev else None)
        for sp in species:
            eng.add_species(sp["coords"], sp["momenta"], sp["weights"], charge=sp["charge"], mass=sp["mass"], still=sp["still"])
        eng.upload("EG_fb", eg0)
        if c["window"]:
            eng.set_window(c["window"][0], staged=c["window"][1])
        eng.make_halfstep(px0=c["px0"], background=any(sp["still"] for sp in species))
        eng.sync()
    return S, species, eg0, c, eng


def small_refrun(fim, S, species, eg0, c):
    """the reference's step sequence (tests/pic_ref.py) on a fimera backend: the per-function drop-in with host
    numpy buffers (e2e) or the CPU oracle (cpu_baseline)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import copy

    from pic_ref import RefRun, RefSpecies

    S = copy.copy(S)
    S.Args = copy.deepcopy(S.Args)  # a moving frame shifts the grid in the solver dictionary
    dev = c["device"]
    sp = [RefSpecies(s["coords"], s["momenta"], s["weights"], charge=s["charge"], mass=s["mass"], still=s["still"],
                     device=(getattr(fim, dev[0]), dev[1]) if (dev and not s["still"]) else None) for s in species]
    run = RefRun(fim, S, sp, background=any(s["still"] for s in species))
    run.EG_fb[:] = eg0
    if c["window"]:
        v, staged = c["window"]
        dt = S.Args["dt"]
        run.window = (0.5 * v * dt, 0.5 * v * dt) if staged else (v * dt, 0.0)
    run.make_halfstep(px0=c["px0"])
    return run


def run_small(a):
    """one GPU, device-resident engine; `value`: every step timed on its own with the L2 flushed in between (the
    whole problem fits the 126 MB L2), `l2_warm`: the K steps in one multi-step call, as a simulation runs them"""
    import torch

    import chimera_b200.fimera as gfim
    from chimera_b200 import _lib

    torch.cuda.set_device(0)
    lib = _lib.load()
    lib.chimera_set_device(0)
    S, species, eg0, c, eng = small_build(a, gfim, engine=True)
    # a side stream that torch's events see too (graph capture is not possible on the legacy default stream)
    side = torch.cuda.Stream()
    eng.use_stream(side.cuda_stream)
    torch.cuda.set_stream(side)
    n = sum(sp["weights"].size for sp in species if not sp["still"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    eng.step(a.warmup)
    eng.sync()
    # (1) L2 flushed between steps; one-step calls (a step's closing gather + push runs inside the next call's fused
    # kernel, the last one in the sync that is timed after the loop)
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    evs = []
    l0 = _lib.kernel_launches()
    for _ in range(a.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.step(1)
        e1.record()
        evs.append((e0, e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.sync()
    e1.record()
    evs.append((e0, e1))
    torch.cuda.synchronize()
    launches = _lib.kernel_launches() - l0
    ms_cold = sum(e0.elapsed_time(e1) for e0, e1 in evs) / a.steps
    # (2) the same K steps as one multi-step call (fused across steps, graph replay between re-binnings)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    eng.step(a.steps)
    eng.sync()
    e1.record()
    torch.cuda.synchronize()
    ms_warm = e0.elapsed_time(e1) / a.steps
    graphs = eng.graph_info()
    clocks = sampler.finish()
    # (3) once more with per-phase events and the contraction profile (graphs off while profiling)
    eng.profile(True)
    eng.timings(reset=True)
    lib.chimera_gemm_profile(1)
    e0.record()
    eng.step(a.steps)
    eng.sync()
    e1.record()
    torch.cuda.synchronize()
    ms_prof = e0.elapsed_time(e1) / a.steps
    phases = eng.timings(reset=True)
    eng.profile(False)
    g_ms, g_fl, g_n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
    lib.chimera_gemm_profile_read(ctypes.byref(g_ms), ctypes.byref(g_fl), ctypes.byref(g_n), 1)
    lib.chimera_gemm_profile(0)
    fp64_peak = dgemm_peak_tflops(torch)
    stages = {k: {"ms_per_call": ms / calls, "share": ms / (ms_prof * a.steps)} for k, (ms, calls) in phases.items()}
    gemm = {"ms_per_launch": g_ms.value / max(g_n.value, 1), "launches_per_step": g_n.value / a.steps,
            "share": g_ms.value / (ms_prof * a.steps), "tflops": g_fl.value / (g_ms.value * 1e-3) / 1e12 if g_ms.value else 0.0}
    out = {
        "metric": METRIC, "value": n / (ms_cold * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_cold, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": SMALL_NAMES[a.config], "particles_total": n, "grid": list(S.shape_sp),
                   "l2": "L2 flushed between timed steps (256 MB write); every step timed on its own with CUDA events"},
        "l2_warm": {"value": n / (ms_warm * 1e-3), "ms_per_step": ms_warm, "step_graphs": graphs[0],
                    "note": "the same K steps as ONE multi-step call (particle work fused across steps, the fused step replayed "
                            "as a CUDA graph where no window moves every step), no flush"},
        "gpu_launches": int(launches), "launches_per_step": launches / a.steps, "clocks": clocks,
        "roofline": {"kernel": "gemm_dmma_k (DHT + mode-coupling contractions)", "bound": "tensor", "achieved": gemm["tflops"],
                     "peak": fp64_peak, "unit": "TFLOP/s", "frac": gemm["tflops"] / fp64_peak, "traffic": None,
                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run", "share_of_step": gemm["share"]},
        "stages": stages, "gemm": gemm, "fp64_peak_tflops": fp64_peak,
        "dht_psatd_ms_per_step": sum(v["ms_per_call"] for k, v in stages.items()
                                     if k in ("fb_in_J", "fb_in_rho", "poisson", "maxwell", "fields_out", "static_fields")),
    }
    eng.close()
    # end to end: the per-function drop-in driven by the reference's make_step sequence -- with HOST numpy buffers (every
    # call copies its arguments in and its results out: `e2e`), and in resident mode (`e2e_resident`: the driver's numpy
    # arrays in CUDA managed memory, no copies)
    if not a.no_e2e:
        import copy

        def drive(resident):
            Sx = S
            if resident:
                gfim.resident(True)
                Sx = copy.deepcopy(S)  # tables built after the switch, as a driver that imports the drop-in first gets them
            try:
                timer = StageTimer(gfim)
                sp2 = [dict(sp, **{k: np.array(sp[k], order="F") for k in ("coords", "momenta", "weights")}) for sp in species]
                run = small_refrun(timer, Sx, sp2, np.array(eg0, order="F"), c)
                for _ in range(2):
                    run.make_step()
                torch.cuda.synchronize()
                lib.chimera_host_traffic(None, None, 1)
                timer.t.clear()
                t = time.perf_counter()
                ne = max(3, min(a.steps, 10))
                for _ in range(ne):
                    run.make_step()
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t) / ne
                h2d, d2h = ctypes.c_longlong(), ctypes.c_longlong()
                lib.chimera_host_traffic(ctypes.byref(h2d), ctypes.byref(d2h), 1)
            finally:
                if resident:
                    gfim.resident(False)
            calls = {k: 1e3 * v / ne for k, v in sorted(timer.t.items(), key=lambda kv: -kv[1])}
            calls["python_statements"] = dt * 1e3 - sum(calls.values())
            return {"value": n / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": ne, "h2d_bytes_per_step": int(h2d.value // ne),
                    "d2h_bytes_per_step": int(d2h.value // ne), "calls_ms": calls,
                    "path": "chimera_b200.fimera per-function drop-in driven by the reference's make_step sequence, " +
                            ("RESIDENT mode: numpy arrays in CUDA managed memory, no staging copies" if resident else
                             "pageable host numpy buffers, every call copies its arguments in and its results out")}

        out["e2e"] = drive(False)
        out["e2e_resident"] = drive(True)
    if not a.no_cpu:
        out["cpu_baseline"] = cpu_small(a, S, species, eg0, c, steps=3, warmup=1)
    print(json.dumps(out))


def cpu_small(a, S, species, eg0, c, steps, warmup):
    fast, cores = load_cpu_oracle()
    timer = StageTimer(fast)
    run = small_refrun(timer, S, species, eg0, c)
    n = sum(sp["weights"].size for sp in species if not sp["still"])
    r = cpu_timed_steps(run, timer, n, steps=steps, warmup=warmup, budget_s=60)
    r.update(unit=UNIT, cores=cores, kind="port",
             sample="the whole workload (%d particles, grid %r), %d warm-up + %d timed RefRun.make_step; g++ -O3 -ffast-math -fopenmp "
                    "build of oracle/chimera_oracle.cpp, OMP threads=%d, deposit chunks=%d as the configuration says"
                    % (n, tuple(S.shape_sp), r["warmup_run"], r["steps_run"], cores, S.Args.get("Xchunked", (1, 0))[0]))
    return r


def run_reference_small(a):
    if int(os.environ.get("RANK", 0)) != 0:
        return
    fast, _ = load_cpu_oracle()
    S, species, eg0, c, _ = small_build(a, fast, engine=False)
    r = cpu_small(a, S, species, eg0, c, steps=a.steps, warmup=1)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": 1, "steps": r["steps_run"],
        "warmup": r["warmup_run"], "steps_requested": a.steps, "warmup_requested": a.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": SMALL_NAMES[a.config]}, "cpu_baseline": r, "gpu_launches": 0,
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port; the Fortran cannot be built here) on the host cores
CPU_STAGES = {  # BASELINE.md section 3 stage list -> the fimera calls that make it up (real and envelope families)
    "push": ("push_coords", "push_velocs", "undul_analytic"), "gather": ("proj_fld", "proj_fld_env"),
    "deposit_J": ("dep_curr_chnk", "dep_curr_env_chnk", "dep_curr", "dep_curr_env"),
    "deposit_rho": ("dep_dens_chnk", "dep_dens_env_chnk", "dep_dens", "dep_dens_env"),
    "rebin": ("chunk_coords_boundaries", "align_data_vec", "align_data_scl", "sortpartsout"),
    "dht_fwd_fft": ("fb_vec_in", "fb_scl_in"), "dht_bwd_fft": ("fb_eb_out",),
    "mode_coupling": ("fb_grad", "fb_graddiv", "fb_rot", "fb_grad_env", "fb_graddiv_env", "fb_rot_env"),
    "psatd": ("maxwell_push_with_spchrg", "maxwell_push_wo_spchrg", "maxwell_init_push", "field_drift"),
    "poisson": ("poiss_corr", "poiss_corr_stat"),
    "elementwise": ("omp_mult_vec", "omp_mult_scl", "omp_add_vec", "eb_correction", "eb_correction_env"),
}


class StageTimer:
    """fimera proxy that accumulates wall time per entry point (the reference driver's calls, unchanged)"""

    def __init__(self, fim):
        self._f, self.t = fim, {}

    def __getattr__(self, name):
        fn = getattr(self._f, name)
        if not callable(fn):
            return fn

        def timed(*args, **kw):
            t0 = time.perf_counter()
            r = fn(*args, **kw)
            self.t[name] = self.t.get(name, 0.0) + time.perf_counter() - t0
            return r
        return timed

    def stages_ms(self, steps):
        out = {k: 1e3 * sum(self.t.get(n, 0.0) for n in names) / steps for k, names in CPU_STAGES.items()}
        return {k: v for k, v in out.items() if v > 0}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def load_cpu_oracle():
    """the -O3 -ffast-math -fopenmp build (reference Makefile:14 flags) with the team size set HERE: torchrun
    exports OMP_NUM_THREADS=1 to its workers, which is not what a CPU run of the reference would use"""
    cores = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    from oracle.fimera import load

    fast = load(fast=True, native=True)  # -march=native when a compiler is on this host
    try:
        cores = int(fast._lib.oracle_set_num_threads(cores))
    except AttributeError:
        cores = int(fast._lib.oracle_num_threads())
    return fast, cores


def cpu_chunks(nx, cores):
    """x-chunks of the CPU run: the chunked deposit runs one thread per chunk (grid_deps_chnk.f90:38-45), so the
    largest chunk count <= cores that divides Nx keeps every core busy (results do not depend on it)"""
    c = min(cores, nx // 2)
    while c > 1 and nx % c:
        c -= 1
    return max(c, 1)


def cpu_timed_steps(run, timer, n_particles_run, steps, warmup, budget_s):
    """`warmup` untimed + up to `steps` timed RefRun.make_step calls (the reference's make_step sequence,
    chimera_main.py:82-92, numpy-side mutations included); the timed count is cut to what fits `budget_s`"""
    for _ in range(warmup):
        t0 = time.perf_counter()
        run.make_step()
        t_one = time.perf_counter() - t0
    steps_run = int(max(1, min(steps, budget_s // max(t_one, 1e-3)))) if warmup else steps
    steps_run = max(steps_run, min(steps, 3))
    timer.t.clear()
    per = []
    for _ in range(steps_run):
        t0 = time.perf_counter()
        run.make_step()
        per.append(time.perf_counter() - t0)
    t_step = float(np.mean(per))
    stages = timer.stages_ms(steps_run)
    stages["python_glue"] = 1e3 * t_step - sum(stages.values())
    return {"value": n_particles_run / t_step, "ms_per_step": 1e3 * t_step, "steps_run": steps_run, "warmup_run": warmup,
            "stages_ms": stages, "ms_per_step_each": [1e3 * v for v in per]}


def cpu_describe(fast, cores, nchnk, what):
    return ("%s; every step = RefRun.make_step (the reference's call sequence) on the FULL grid and particle set, no "
            "extrapolation; g++ -O3 -ffast-math -fopenmp build of oracle/chimera_oracle.cpp (%s), OMP threads=%d, "
            "deposit chunks=%d (one thread per chunk as grid_deps_chnk.f90:38-45); gfortran/FFTW3 absent so the Fortran "
            "itself cannot be built; the re-binning step (every 11th, numpy argsort) is timed separately as rebin_s"
            % (what, "-march=native, built on this host" if fast.build_name == "liboracle_native.so" else "-march=x86-64-v3",
               cores, nchnk))


def cpu_baseline(a, S, state, bck):
    """cpu_baseline of our arm (N=1): the engine's own state (particles, EG_fb, gradRho, background) downloaded
    to the host, re-binned into `cores` chunks, then 1 warm-up + 2 timed full steps on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import copy

    from chimera_b200.solver_setup import SolverSetup
    from pic_ref import RefRun, RefSpecies

    fast, cores = load_cpu_oracle()
    timer = StageTimer(fast)
    x, xh, p, w, eg, g = state
    S2 = copy.copy(S)
    nchnk = cpu_chunks(a.nx, cores)
    S2.Args = dict(S.Args, Xchunked=(nchnk, S.Args["Xchunked"][1]))
    sp = RefSpecies.__new__(RefSpecies)
    sp.coords, sp.momenta, sp.weights, sp.coords_halfstep = x, p, w, xh
    sp.push_fact, sp.still, sp.devices, sp.chunks = -2 * np.pi, False, [], None
    sp.EB = np.zeros((6, 0), order="F")
    run = RefRun(timer, S2, [sp], sort_every=0)
    run.Bck, run.EG_fb, run.g_nxt = bck, eg, g
    t0 = time.perf_counter()
    run.chunk_and_damp(sp, "stag")
    rebin_s = time.perf_counter() - t0
    n = sp.weights.shape[0]
    r = cpu_timed_steps(run, timer, n, steps=2, warmup=1, budget_s=60)
    r.update(unit=UNIT, cores=cores, kind="port", rebin_s=rebin_s,
             sample=cpu_describe(fast, cores, nchnk, "%d particles (all of them) on the %dx%dx%d grid, 1 warm-up + %d timed steps"
                                 % (n, a.nx, a.nr + 1, a.modes, r["steps_run"])))
    return r


def run_reference(a):
    """--impl reference: the stated workload in full (whole grid, every particle of the N-GPU job: one species per
    GPU rank, same seeds as the GPU arm) through the reference's make_step sequence on all host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from concurrent.futures import ThreadPoolExecutor

    from chimera_b200 import synthetic
    from chimera_b200.solver_setup import SolverSetup
    from pic_ref import RefRun, RefSpecies

    world = int(os.environ.get("WORLD_SIZE", a.gpus))
    fast, cores = load_cpu_oracle()
    timer = StageTimer(fast)
    nchnk = cpu_chunks(a.nx, cores)
    S = SolverSetup(synthetic.lwfa_solver_config(nx=a.nx, nr=a.nr, modes=a.modes, chunks=nchnk))
    n_gpu = n_particles(a)
    shares = world
    try:  # x, x_half, p, w, per-particle EB and the generator's temporaries: ~260 B per particle at the peak
        import psutil

        avail = psutil.virtual_memory().available
        while shares > 1 and shares * n_gpu * 260.0 > 0.8 * avail:
            shares -= 1
    except Exception:
        pass
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=min(shares, 8)) as ex:
        parts = list(ex.map(lambda r: synthetic.plasma_fixed_cell(S.Args, cell=CELL[a.ppc], seed=20260101 + r, xp=np), range(shares)))
    species = [RefSpecies(x, p, w) for x, p, w in parts]
    del parts
    # still ions coincide with the electrons at t=0: BckGrndRho = -rho_e(0), as build_problem does on the GPU
    run = RefRun(timer, S, species)
    for s in species:
        run.chunk_and_damp(s, "stag")
    run.Rho[:] = 0.0
    for s in species:
        run.Rho = run._dep("dens", run.Rho, s, s.coords)
    run.Bck = -run.Rho
    run.EG_fb[:] = synthetic.laser_seed(S, fast)
    timer.t.clear()
    t1 = time.perf_counter()
    run.make_halfstep()
    rebin_s = sum(timer.t.get(k, 0.0) for k in CPU_STAGES["rebin"])
    setup_s = (t1 - t0, time.perf_counter() - t1)
    n_run = sum(s.weights.shape[0] for s in species)
    r = cpu_timed_steps(run, timer, n_run, steps=a.steps, warmup=max(1, min(a.warmup, 1)), budget_s=150)
    what = ("%d particles = %d of the %d GPU shares of %d, on the %dx%dx%d grid; %d warm-up + %d timed steps of the %d/%d requested "
            "(cut to a 150 s budget)" % (n_run, shares, world, n_gpu, a.nx, a.nr + 1, a.modes, r["warmup_run"], r["steps_run"],
                                         a.warmup, a.steps))
    r.update(unit=UNIT, cores=cores, kind="port", rebin_s=rebin_s, setup_s=setup_s, sample=cpu_describe(fast, cores, nchnk, what))
    out = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": r["steps_run"],
        "warmup": r["warmup_run"], "steps_requested": a.steps, "warmup_requested": a.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "particles_total": n_run}, "cpu_baseline": r, "gpu_launches": 0,
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


if __name__ == "__main__":
    args = parse()
    if args.config != "c3":
        (run_reference_small if args.impl == "reference" else run_small)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
