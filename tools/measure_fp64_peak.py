"""Measure the FP64 GEMM denominators on this GPU (MEASURED_PEAKS.json has no FP64 figure):
cuBLAS DGEMM 8192^3 via torch.matmul (burst: best of 10; sustained: back to back for ~3 s), and
this library's DMMA contraction kernel on the DHT shapes.  Writes gpurun_out/fp64_peak.json."""
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chimera_b200 import _lib  # noqa: E402


def cublas_dgemm(n=8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2 * n ** 3 / (best * 1e-3) / 1e12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); k = 0
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(5):
            a @ b
        k += 5
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sustained = 2 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return burst, sustained


def ours(nkx, K, N, batch, iters=20):
    lib = _lib.load()
    ms = ctypes.c_double(0)
    rc = lib.chimera_bench_gemm(ctypes.c_longlong(nkx), ctypes.c_longlong(K), ctypes.c_longlong(N), batch, iters, ctypes.byref(ms))
    if rc:
        raise RuntimeError(lib.chimera_last_error().decode())
    flop = 4.0 * nkx * K * N * batch
    return ms.value, flop / (ms.value * 1e-3) / 1e12


if __name__ == "__main__":
    out = {"gpu": torch.cuda.get_device_name(0)}
    out["cublas_dgemm_8192_tflops_burst"], out["cublas_dgemm_8192_tflops_sustained"] = cublas_dgemm()
    out["dmma_kernel"] = []
    for (nkx, K, N, batch) in [(4096, 512, 512, 9), (4096, 512, 512, 18), (1024, 256, 256, 9), (16384, 1024, 1024, 3),
                               (4096, 2048, 2048, 3), (1272, 64, 64, 12), (304, 300, 300, 12)]:
        ms, tf = ours(nkx, K, N, batch)
        out["dmma_kernel"].append({"nkx": nkx, "K": K, "N": N, "batch": batch, "ms": ms, "tflops": tf})
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/fp64_peak.json", "w"), indent=1)
    print(json.dumps(out, indent=1))
