"""Staged pageable copies (csrc/staging.cu) at the thread count / chunk size given by CHIMERA_STAGE_THREADS /
CHIMERA_STAGE_CHUNK_MB: one H2D of 2.4 GB + one D2H of 1.2 GB through align_data_scl."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import chimera_b200.fimera as gfim
m = 150_000_000
z, idx = np.ones(m), np.arange(m, dtype=np.int64)
gfim.align_data_scl(z[:1000000].copy(), idx[:1000000])
best = 1e9
for _ in range(3):
    t = time.perf_counter(); gfim.align_data_scl(z, idx); best = min(best, time.perf_counter() - t)
print("threads=%s chunk=%s MB: %.1f ms = %.1f GB/s" % (os.environ.get("CHIMERA_STAGE_THREADS", "default"),
      os.environ.get("CHIMERA_STAGE_CHUNK_MB", "16"), best * 1e3, 3 * m * 8 / 1e9 / best))
