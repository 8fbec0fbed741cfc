import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from util import SETUPS, rel_l2
from oracle import fimera as ofim
import chimera_b200.fimera as gfim
import test_gpu_engine as T

for und in (None, dict(a0=0.3, **{"lambda": 1.3}, X0=-1.0, Lx=9.0)):
    S, ref, eng = T.build_pair(ofim, "env_m1", 11, undulator=und)
    ref.make_halfstep(); eng.make_halfstep()
    for n in ("J", "J_fb", "EG_fb", "B_fb", "EB"):
        print(und is not None, n, rel_l2(eng.download(n), getattr(ref, n)))
    x, xh, p, w = eng.particles(0)
    perm = T.match(ref.sp[0].weights, w)
    print("momenta", rel_l2(p[:, perm], ref.sp[0].momenta), "coords", rel_l2(x[:, perm], ref.sp[0].coords))
    # gather alone, both backends, on the oracle's EB
    s = ref.sp[0]
    a = S.Args
    eo = ofim.proj_fld_env(s.coords, s.weights, ref.EB, np.zeros((6, s.coords.shape[1]), order="F"), a["leftX"], *a["DepProj"])
    eg = gfim.proj_fld_env(s.coords, s.weights, ref.EB, np.zeros((6, s.coords.shape[1]), order="F"), a["leftX"], *a["DepProj"])
    print("proj_fld_env alone", rel_l2(eg, eo), np.abs(eo).max(), np.abs(ref.EB).max())
    d = np.abs(p[:, perm] - s.momenta)
    print("max abs dp", d.max(), "at", np.unravel_index(d.argmax(), d.shape), "p scale", np.abs(s.momenta).max())
    eng.close()
