"""SR post-processing throughput: the CUDA path through the fimera-compatible C-ABI call
(host buffers, copies inside the timed region) next to the CPU oracle port (-O3 -ffast-math -fopenmp, all host
cores) on a bounded particle sample.  Unit: integrand terms/s, one term = one (particle, time step, frequency,
pixel) evaluation of A e^{i omega phi}.  Writes gpurun_out/sr_bench.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import chimera_b200.fimera as gfim  # noqa: E402
from oracle.fimera import load  # noqa: E402
from test_sr import far_grid, near_args, tracks  # noqa: E402

ofast = load(fast=True)
out = {"unit": "terms/s", "cases": []}
NT, NP, NOM = 2048, 2048, 256
x, mp, mn, w, dt = tracks(NT, NP, 1)
for mode, comp in (("far", 0), ("far", 1), ("near", 0), ("nearcirc", 0)):
    if mode == "far":
        g = far_grid(NOM, 16, 8)
        n1, n2 = 16, 8
        name = "sr_calc_far_tot" if comp == 0 else "sr_calc_far_comp"
        mk = lambda f, xs, a, b, ws: getattr(f, name)(np.zeros((NOM, n1, n2), order="F"), xs, a, b, ws,  # noqa: E731
                                                      *([comp] if comp else []), dt, *g)
    else:
        n1, n2 = 16, 8
        g = near_args(mode == "nearcirc", NOM, n1, n2)
        name = "sr_calc_%s_tot" % mode
        mk = lambda f, xs, a, b, ws: getattr(f, name)(np.zeros((NOM, n1, n2), order="F"), xs, b, ws, dt, *g)  # noqa: E731
    mk(gfim, x[:, :, :64], mp[:, :, :64], mn[:, :, :64], w[:64])  # warm-up
    t = time.perf_counter()
    sg = mk(gfim, x, mp, mn, w)
    tg = time.perf_counter() - t
    ns = 2 * (os.cpu_count() or 1)
    sl = [np.asfortranarray(a[:, :, :ns]) for a in (x, mp, mn)]
    t = time.perf_counter()
    sc = mk(ofast, sl[0], sl[1], sl[2], w[:ns])
    tc = time.perf_counter() - t
    sgs = mk(gfim, sl[0], sl[1], sl[2], w[:ns])
    terms = float(NT) * NOM * n1 * n2
    out["cases"].append({"call": name, "nt": NT, "np": NP, "nom": NOM, "pixels": n1 * n2,
                         "gpu_terms_per_s": terms * NP / tg, "gpu_s": tg,
                         "cpu_terms_per_s": terms * ns / tc, "cpu_s": tc, "cpu_sample_particles": ns,
                         "cpu_cores": os.cpu_count(), "speedup": (terms * NP / tg) / (terms * ns / tc),
                         "rel_l2_vs_cpu_sample": float(np.linalg.norm(sgs - sc) / np.linalg.norm(sc))})
    print(out["cases"][-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sr_bench.json"), "w"), indent=1)
