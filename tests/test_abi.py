"""CPU: the C-ABI boundary.  libchimera_b200.so loads without a GPU, exports every symbol that
include/chimera_b200.h declares, and its compute entry points fail loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "chimera_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return sorted(set(re.findall(r"\b(chimera_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_fimera_hot_path():
    names = declared_symbols()
    # one entry point per Fortran subroutine the reference driver calls on the path
    for n in ("push_velocs push_coords genparts sortpartsout chunk_coords_boundaries align_data_vec align_data_scl "
              "dep_curr dep_dens proj_fld eb_correction dep_curr_chnk dep_dens_chnk dep_curr_env dep_dens_env proj_fld_env "
              "eb_correction_env dep_curr_env_chnk dep_dens_env_chnk fb_vec_in fb_scl_in fb_vec_out fb_scl_out fb_eb_out "
              "fb_filtr fb_rot fb_grad fb_div fb_graddiv fb_rot_env fb_grad_env fb_div_env fb_graddiv_env "
              "maxwell_push_with_spchrg maxwell_push_wo_spchrg maxwell_init_push poiss_corr poiss_corr_stat field_drift "
              "omp_mult_vec omp_mult_scl omp_add_vec omp_add_scl undul_analytic undul_analytic_taper undul_mapped "
              "undul_mapped_tap planewave gaussbeam").split():
        assert "chimera_" + n in names, n


def test_library_exports_every_declared_symbol():
    from chimera_b200 import _lib

    lib = _lib.load()
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.chimera_version().decode()


def test_shim_covers_every_per_function_entry_point():
    import chimera_b200.fimera as f

    skip = ("last_error", "version", "device_count", "set_device", "sync", "kernel_launches", "host_traffic", "bench_gemm",
            "gemm_profile", "gemm_profile_read", "fused_profile", "fused_profile_read", "host_register", "host_unregister")
    for sym in declared_symbols():
        name = sym[len("chimera_"):]
        if name.startswith("engine_") or name in skip:
            continue
        assert hasattr(f, name), name


def test_compute_fails_loudly_without_a_gpu():
    import chimera_b200.fimera as f

    if f.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(f.error):
        f.push_velocs(np.zeros((3, 4), order="F"), np.zeros((6, 4), order="F"), 0.1)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "chimera_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "liboracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, fn
