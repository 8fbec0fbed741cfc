"""Contraction kernel alone at the shapes the step launches (chimera_bench_gemm: back-to-back launches, CUDA events).
   python tools/gemm_sweep.py [tag]  ->  gpurun_out/gemm_sweep_<tag>.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from measure_fp64_peak import ours  # noqa: E402

SHAPES = [  # (nkx, K, N, batch)
    (304, 300, 300, 1), (304, 300, 300, 3), (304, 300, 300, 4), (304, 300, 300, 12),   # space-charge demo (C2)
    (528, 64, 64, 4), (528, 64, 64, 12), (120, 84, 84, 3),                             # LPA / FEL demo (C1)
    (512, 512, 512, 3), (512, 512, 512, 12),                                           # kx slab of an 8-GPU run
    (4096, 512, 512, 3), (4096, 512, 512, 6), (4096, 512, 512, 18),                    # LWFA bench (C3)
]
if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "run"
    out = []
    for (nkx, K, N, b) in SHAPES:
        ms, tf = ours(nkx, K, N, b, iters=50)
        out.append({"nkx": nkx, "K": K, "N": N, "batch": b, "us": ms * 1e3, "tflops": tf})
        print("%5d x %4d x %4d  batch %2d   %8.1f us   %6.2f TF/s" % (nkx, K, N, b, ms * 1e3, tf), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/gemm_sweep_%s.json" % tag, "w"), indent=1)
