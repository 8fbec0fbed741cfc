"""Golden fixture for the laser injection: the reference's unmodified ``Solver.add_gauss_beam`` (moduls/solvers.py:555-603)
with the CPU oracle as ``chimera.moduls.fimera`` (BUILD CONTAINER ONLY: needs /root/reference) -> tests/golden/laser.npz.
Pins ``SolverSetup.add_gauss_beam`` (chimera_b200/solver_setup.py).    python tools/gen_golden_laser.py"""
import sys, os, copy, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, ROOT+'/tests', ROOT+'/tools']
import ref_driver
from oracle import fimera as ofim
R = ref_driver.install(ofim)
from util import SETUPS
out={}
for name, laser in (("real_m2", dict(a0=3.0, k0=1.0, x0=-1.2, x_foc=2.5, Lx=0.5, LR=1.0)),
                    ("env_m1", dict(a0=0.15, k0=SETUPS["env_m1"]["KxShift"], x0=-0.6, x_foc=25.0, Lx=1.2, LR=9.0)),
                    ("env_m3", dict(a0=0.2, k0=30.0, x0=0.3, x_foc=-4.0, Lx=0.8, LR=2.0))):
    np.random.seed(1)
    sol = R.Solver(copy.deepcopy(SETUPS[name]))
    sol.add_gauss_beam(dict(laser))
    out[name+"_EG_fb"] = np.array(sol.Data["EG_fb"], order="F")
    out[name+"_laser"] = np.array([laser[k] for k in ("a0","k0","x0","x_foc","Lx","LR")])
    print(name, np.abs(out[name+"_EG_fb"]).max())
np.savez_compressed(ROOT+"/tests/golden/laser.npz", **out)
print(os.path.getsize(ROOT+"/tests/golden/laser.npz")/1e3, "kB")
