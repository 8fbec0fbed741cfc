"""Golden fixture for the synchrotron-radiation path, recorded from the reference's UNMODIFIED ``SR`` class
(moduls/SR.py) with the CPU oracle installed as ``chimera.moduls.fimera`` (BUILD CONTAINER ONLY: needs
/root/reference; nothing is copied).  The class builds the frequency / angle / screen grids and the argument
lists (``Args['DepFact']``, SR.py:71-73,97-98,124-126), stores the tracks (``init_track`` / ``add_track``,
SR.py:128-151: far field keeps momenta before and after the push and full-step coordinates, near field keeps
half-step coordinates) and calls ``sr_calc_*`` (SR.py:165-215); its numpy post-processing
(``get_energy_spectrum``, ``get_energy``) is recorded as well.  As for the PIC fixtures, the call sequence and
the Python-side arithmetic are the reference's own, the kernels underneath are the oracle restatement.

  python tools/gen_golden_sr.py            # rewrites tests/golden/sr.npz
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

GAMMA = 20.0
MODES = {
    "far": {"Mode": "far", "Grid": [(5.0, 2.5 * GAMMA ** 2 / 3.0), (0.0, 2.0 / GAMMA), (0.0, 2 * np.pi), (24, 4, 3)]},
    "near": {"Mode": "near", "Grid": [(5.0, 300.0), (-3.0, 3.0), (-2.0, 2.0), 40.0, (24, 3, 4)]},
    "near-circ": {"Mode": "near-circ", "Grid": [(5.0, 300.0), (0.0, 3.0), (0.0, 2 * np.pi), 40.0, (24, 3, 4)],
                  "Features": ("WavelengthGrid",)},
}


def main():
    import ref_driver
    from oracle import fimera as ofim
    from test_sr import tracks

    ref_driver.install(ofim)
    from chimera.moduls.SR import SR

    nt, n = 90, 6
    x, mp, mn, w, dt = tracks(nt, n, 77, gamma=GAMMA)
    out = {"coords": x, "momenta_prv": mp, "momenta_nxt": mn, "weights": w, "dt": np.array(dt)}
    for name, args in MODES.items():
        sr = SR(dict(args, TimeStep=dt))
        beam = types.SimpleNamespace(Data={"coords": x[:, 0].copy(order="F"), "coords_halfstep": x[:, 0].copy(order="F"),
                                           "momenta": mp[:, 0].copy(order="F"), "weights": w.copy()})
        sr.init_track(nt, beam)
        for it in range(nt):  # what the user's loop does after every make_step
            beam.Data["coords"] = x[:, it].copy(order="F")
            beam.Data["coords_halfstep"] = x[:, it].copy(order="F")
            beam.Data["momenta"] = mn[:, it].copy(order="F")
            if it == 0:
                beam.Data["momenta_prv"] = mp[:, 0].copy(order="F")
            sr.add_track(beam)
        key = name.replace("-", "")
        for i, v in enumerate(sr.Args["DepFact"]):
            out["%s_depfact%d" % (key, i)] = np.array(v)
        sr.calculate_spectrum(comp="all")
        out[key + "_rad_all"] = sr.Data["Rad"].copy(order="F")
        out[key + "_energy_spectrum"] = sr.get_energy_spectrum(chim_units=True, lambda0_um=0.8)
        out[key + "_energy"] = np.array(sr.get_energy(chim_units=True, lambda0_um=0.8))
        sr.Data["Rad"][:] = 0.0
        sr.calculate_spectrum(comp="y")
        out[key + "_rad_y"] = sr.Data["Rad"].copy(order="F")
        print("%-10s |Rad| %.6e  energy %.6e" % (name, np.linalg.norm(out[key + "_rad_all"].ravel()), float(out[key + "_energy"])))
    path = os.path.join(ROOT, "tests", "golden", "sr.npz")
    np.savez_compressed(path, **out)
    print("%s %.1f kB" % (path, os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
