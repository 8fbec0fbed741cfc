"""Write PARITY.md: for every `fimera` entry point, where it is restated and by which tests it is checked.
Derived from the sources (grep), so it cannot drift from them:  python tools/parity_matrix.py"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fimera as ofim  # noqa: E402

REF = {  # reference file:line of each subroutine (SURVEY.md section 8a / include/chimera_b200.h)
    "push_velocs": "particle_tools.f90:18", "push_coords": "particle_tools.f90:58", "genparts": "particle_tools.f90:84",
    "sortpartsout": "particle_tools.f90:130", "chunk_coords_boundaries": "particle_tools.f90:155",
    "align_data_vec": "particle_tools.f90:270", "align_data_scl": "particle_tools.f90:298", "sortoutghosts": "particle_tools.f90:326",
    "dep_curr": "grid_deps.f90:18", "dep_dens": "grid_deps.f90:89", "proj_fld": "grid_deps.f90:149", "eb_correction": "grid_deps.f90:219",
    "dep_curr_chnk": "grid_deps_chnk.f90:18", "dep_dens_chnk": "grid_deps_chnk.f90:132",
    "dep_curr_env": "grid_deps_env.f90:18", "dep_dens_env": "grid_deps_env.f90:96", "proj_fld_env": "grid_deps_env.f90:164",
    "eb_correction_env": "grid_deps_env.f90:240", "dep_curr_env_chnk": "grid_deps_env_chnk.f90:18", "dep_dens_env_chnk": "grid_deps_env_chnk.f90:140",
    "fb_vec_in": "fb_io.f90:18", "fb_scl_in": "fb_io.f90:61", "fb_vec_out": "fb_io.f90:100", "fb_scl_out": "fb_io.f90:142",
    "fb_eb_out": "fb_io.f90:182", "fb_filtr": "fb_io.f90:230",
    "fb_rot": "fb_math.f90:18", "fb_grad": "fb_math.f90:96", "fb_div": "fb_math.f90:151", "fb_graddiv": "fb_math.f90:201",
    "fb_grad_env": "fb_math_env.f90:18", "fb_div_env": "fb_math_env.f90:63", "fb_rot_env": "fb_math_env.f90:106", "fb_graddiv_env": "fb_math_env.f90:164",
    "maxwell_push_with_spchrg": "maxwell_solvers.f90:18", "maxwell_push_wo_spchrg": "maxwell_solvers.f90:62",
    "maxwell_init_push": "maxwell_solvers.f90:98", "poiss_corr": "maxwell_solvers.f90:131", "poiss_corr_stat": "maxwell_solvers.f90:166",
    "field_drift": "maxwell_solvers.f90:199", "omp_mult_vec": "maxwell_solvers.f90:228", "omp_mult_scl": "maxwell_solvers.f90:252",
    "omp_add_vec": "maxwell_solvers.f90:274", "omp_add_scl": "maxwell_solvers.f90:298",
    "undul_mapped": "devices.f90:18", "undul_mapped_tap": "devices.f90:64", "undul_analytic_taper": "devices.f90:117",
    "undul_analytic": "devices.f90:162", "planewave": "devices.f90:205", "gaussbeam": "devices.f90:254",
    "sr_calc_far_tot": "SR.f90:18", "sr_calc_far_comp": "SR.f90:139", "sr_calc_near_comp": "SR.f90:256", "sr_calc_near_tot": "SR.f90:352",
    "sr_calc_nearcirc_comp": "SR.f90:449", "sr_calc_nearcirc_tot": "SR.f90:546",
    "intens_profo": "utils.f90:18", "density_2x": "utils.f90:210",
}
NP_ALIAS = {"fb_vec_in": "fb_in", "fb_scl_in": "fb_in", "fb_vec_out": "fb_out", "fb_scl_out": "fb_out", "align_data_vec": "align_data",
            "align_data_scl": "align_data", "sr_calc_far_tot": "sr_calc_far", "sr_calc_far_comp": "sr_calc_far",
            "sr_calc_near_tot": "sr_calc_near", "sr_calc_near_comp": "sr_calc_near", "sr_calc_nearcirc_tot": "sr_calc_near",
            "sr_calc_nearcirc_comp": "sr_calc_near", "omp_mult_vec": "(inline numpy)", "omp_mult_scl": "(inline numpy)",
            "omp_add_vec": "(inline numpy)", "omp_add_scl": "(inline numpy)"}


def read(p):
    return open(os.path.join(ROOT, p)).read()


def main():
    tests = {f: read(os.path.join("tests", f)) for f in sorted(os.listdir(os.path.join(ROOT, "tests"))) if f.endswith(".py")}
    npref = read("oracle/np_ref.py")
    cpp = read("oracle/chimera_oracle.cpp") + read("oracle/sr_utils_oracle.cpp")
    hdr = read("include/chimera_b200.h")
    cu = "".join(read(os.path.join("chimera_b200/csrc", f)) for f in os.listdir(os.path.join(ROOT, "chimera_b200/csrc")) if f.endswith(".cu"))
    gpu_files = [f for f in tests if f.startswith("test_gpu") or f in ("test_golden.py", "test_sr.py")]
    rows = []
    import ctypes

    lib = ctypes.CDLL(os.path.join(ROOT, "chimera_b200", "libchimera_b200.so"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import f90_cases

    f90_funcs = {fn for _, fn, _ in f90_cases.cases(ofim)}
    for name in ofim.API_NAMES:
        # tests spell the variants literally or build them as base + "_env" / "_chnk" (tests/test_gpu_parity.py:126-250,
        # tests/test_np_ref.py, tests/pic_ref.py:79): accept the base name next to a suffix expression as well
        base = name.replace("_env", "").replace("_chnk", "")
        alts = [r"\b%s\b" % re.escape(name)]
        if base != name:
            alts += [r'"%s" \+ ' % re.escape(base), r'"%s%%s' % re.escape(base)]
        if name.startswith("sr_calc_near"):
            alts.append(r'"sr_calc_near" \+ ')
        pat = re.compile("|".join(alts))
        npname = NP_ALIAS.get(name, name)
        has_np = npname.startswith("(") or re.search(r"^def %s\(" % re.escape(npname), npref, re.M)
        cpu_t = [f for f, t in tests.items() if not f.startswith("test_gpu") and f.startswith("test_") and pat.search(t)]
        # the step-level tests reach most kernels through the sequence in tests/pic_ref.py
        via_seq = bool(pat.search(tests.get("pic_ref.py", ""))) or name.startswith(("dep_curr", "dep_dens"))
        gpu_t = [f for f in gpu_files if pat.search(tests[f]) and "pytest.mark.gpu" in tests[f] or (f.startswith("test_gpu") and pat.search(tests[f]))]
        # executed reference Fortran: the vectors of tests/golden/f90_kernels.npz are keyed "<case>|<k>", the case table
        # (tests/f90_cases.py) names the fimera function of every case
        f90 = "yes" if name in f90_funcs else "—"
        rows.append("| `%s` | %s | %s | %s | %s | %s | %s | %s |" % (
            name, REF.get(name, "?"), f90,
            "yes" if ("oracle_%s" % name) in cpp else "NO",
            ("`%s`" % npname if not npname.startswith("(") else npname) if has_np else "—",
            "yes" if ("chimera_%s(" % name) in hdr and hasattr(lib, "chimera_%s" % name) else "NO",
            ", ".join(sorted(set(t.replace("test_", "").replace(".py", "") for t in cpu_t))) or "—",
            (", ".join(sorted(set(t.replace("test_", "").replace(".py", "") for t in gpu_t))) or "—") + (" + step sequence" if via_seq else "")))
    out = ["# PARITY — coverage per `fimera` entry point", "",
           "Generated by `tools/parity_matrix.py` from the sources. Columns: reference subroutine; whether golden vectors from",
           "the reference's OWN Fortran exist for it (executed by `oracle/f90py.py`, `tests/golden/f90_kernels.npz`, checked",
           "against the oracle and the CUDA library by `tests/test_f90_golden.py`); C++ oracle restatement",
           "(`oracle/*.cpp`); independent numpy restatement (`oracle/np_ref.py`); CUDA entry point (`include/chimera_b200.h` +",
           "`chimera_b200/csrc`); CPU tests that exercise it (oracle vs numpy, known answers, golden fixtures, shim); GPU tests",
           "(CUDA vs oracle / fixtures). \"step sequence\": also reached by every engine / drop-in step test through",
           "`tests/pic_ref.py`, the restatement of `ChimeraRun.make_halfstep/make_step/frame_act` that the golden fixtures",
           "recorded from the reference's own driver pin.", "",
           "| entry point | reference | executed Fortran vectors | C++ oracle | numpy restatement | CUDA | CPU tests | GPU tests |",
           "|---|---|---|---|---|---|---|---|"]
    out += rows
    out += ["", "%d entry points; the reference's Python uses 50 of them (SURVEY.md section 8b)." % len(rows), ""]
    open(os.path.join(ROOT, "PARITY.md"), "w").write("\n".join(out))
    print("\n".join(out[:14] + rows[:6]))
    missing = [r for r in rows if "| NO |" in r or "| — | yes" in r or r.rstrip().endswith("| — |")]
    print("rows without a numpy restatement or missing pieces:", len(missing))
    for r in missing:
        print(r[:120])


if __name__ == "__main__":
    main()
