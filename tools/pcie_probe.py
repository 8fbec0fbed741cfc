"""Measure host<->device copy bandwidth with page-locked numpy memory (cudaHostRegister) and with torch pinned
memory, one direction and both directions at once: the ceiling of bench.py's end-to-end leg."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes
from chimera_b200 import _lib
lib = _lib.load()
n = 300_000_000  # 2.4 GB
dev_a = torch.empty(n, dtype=torch.float64, device="cuda")
dev_b = torch.empty(n, dtype=torch.float64, device="cuda")
def bw(label, h_in, h_out):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize(); t = time.perf_counter()
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1): dev_a.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2): h_out.copy_(dev_b, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        gb = n * 8 / 1e9 * (2 if mode == "both" else 1)
        print("%-22s %-5s %6.1f GB/s (%.1f ms)" % (label, mode, gb / dt, dt * 1e3))
a = torch.empty(n, dtype=torch.float64).pin_memory(); b = torch.empty(n, dtype=torch.float64).pin_memory()
a.fill_(1.0); b.fill_(2.0)
bw("torch pinned", a, b); bw("torch pinned (2nd)", a, b)
x = np.ones(n); y = np.ones(n)
t = time.perf_counter()
lib.chimera_host_register(ctypes.c_void_p(x.ctypes.data), ctypes.c_longlong(x.nbytes)); lib.chimera_host_register(ctypes.c_void_p(y.ctypes.data), ctypes.c_longlong(y.nbytes))
print("cudaHostRegister 2 x 2.4 GB: %.1f ms" % ((time.perf_counter() - t) * 1e3))
bw("numpy registered", torch.from_numpy(x), torch.from_numpy(y)); bw("numpy registered (2nd)", torch.from_numpy(x), torch.from_numpy(y))
# pageable numpy memory: what a per-function drop-in call sees (plain copies) and what csrc/staging.cu makes of it
# (align_data_scl with the identity permutation = one staged H2D + one staged D2H of the array, plus a tiny kernel)
import chimera_b200.fimera as gfim
m = 150_000_000  # 1.2 GB
z = np.ones(m)
torch.cuda.synchronize(); t = time.perf_counter(); dev_a[:m].copy_(torch.from_numpy(z)); torch.cuda.synchronize()
print("pageable plain        h2d   %6.1f GB/s" % (m * 8 / 1e9 / (time.perf_counter() - t)))
torch.cuda.synchronize(); t = time.perf_counter(); torch.from_numpy(z).copy_(dev_a[:m]); torch.cuda.synchronize()
print("pageable plain        d2h   %6.1f GB/s" % (m * 8 / 1e9 / (time.perf_counter() - t)))
idx = np.arange(m, dtype=np.int64)
gfim.align_data_scl(z[:1000].copy(), idx[:1000])
t = time.perf_counter(); gfim.align_data_scl(z, idx); dt = time.perf_counter() - t
print("pageable staged (csrc/staging.cu) h2d 2.4 GB + d2h 1.2 GB in %.1f ms = %.1f GB/s" % (dt * 1e3, 3 * m * 8 / 1e9 / dt))
