"""CPU: the f2py calling conventions reproduced by chimera_b200/f2py_shim.py,
exercised on the oracle backend (same shim code as the CUDA backend)."""
import numpy as np
import pytest

from util import crandn, setup


def test_inout_arrays_are_modified_in_place_and_returned(ofim):
    p = np.asfortranarray(np.ones((3, 5)))
    f = np.asfortranarray(np.ones((6, 5)))
    q = ofim.push_velocs(p, f, 0.1)
    assert q is p and not np.allclose(p, 1.0)


def test_c_ordered_input_is_converted_to_a_fortran_copy(ofim):
    """an array born empty turns C-ordered on its first resize (species.py:234); f2py converts it and the
    driver adopts the returned copy"""
    p = np.ones((3, 5))  # C order
    q = ofim.push_velocs(p, np.ones((6, 5), order="F"), 0.1)
    assert q is not p and q.flags.f_contiguous and np.allclose(p, 1.0)


def test_multiple_outputs_come_back_in_dummy_argument_order(ofim):
    S = setup("real_m2")
    a = S.Args
    x = np.asfortranarray(np.vstack((np.linspace(a["leftX"], a["rightX"], 50), np.full(50, 0.3), np.zeros(50))))
    dom = np.asfortranarray([a["leftX"], a["rightX"], 0.0, a["Rgrid"].max() ** 2])
    ids, chunks, go_out = ofim.chunk_coords_boundaries(x, dom, a["Xgrid"], 4)
    assert ids.dtype == np.int8 and ids.shape == (50,) and chunks.dtype == np.int32 and chunks.shape == (5,)
    assert isinstance(go_out, int) and chunks[-1] + go_out == 50
    xn, xc = ofim.push_coords(x, np.zeros_like(x), np.zeros_like(x), 0.1)
    assert xn is x and xc.shape == x.shape
    idx, n = ofim.sortpartsout(x, dom)
    assert idx.dtype == np.int32 and isinstance(n, int)


def test_returned_particle_arrays_own_their_data(ofim):
    """the driver resizes what fimera returns (species.py:394-398)"""
    dat = np.asfortranarray(np.arange(30.0).reshape(3, 10))
    out = ofim.align_data_vec(dat, np.array([4, 2, 7], dtype=np.int32))  # int32 index is cast like f2py does
    assert out.flags.owndata
    out.resize((3, 3), refcheck=False)
    assert np.array_equal(out[0], [4.0, 2.0, 7.0]) and np.array_equal(out[2], [24.0, 22.0, 27.0])


def test_shape_mismatch_raises_module_error(ofim):
    with pytest.raises(ofim.error):
        ofim.push_velocs(np.zeros((3, 5), order="F"), np.zeros((6, 4), order="F"), 0.1)
    S = setup("real_m2")
    with pytest.raises(ofim.error):
        ofim.fb_vec_in(S.zeros_fb(3), S.zeros_sp(3)[:-1], 0.0, S.Args["kx"], S.Args["In"])
    with pytest.raises(ofim.error):  # the envelope variants need an odd number of mode slots
        ofim.eb_correction_env(np.zeros((8, 4, 2, 6), dtype=complex, order="F"))


def test_slices_of_a_parent_array_alias_it(ofim):
    """EG_fb[:, :, :, 3:] is passed as an input and EG_fb[:, :, :, :3] as an in/out (solvers.py:548, 621-632)"""
    S = setup("real_m2")
    eg = crandn(np.random.default_rng(1), S.shape_fb + (6,))
    view = eg[:, :, :, :3]
    assert view.flags.f_contiguous
    before = eg[:, :, :, 3:].copy()
    out = ofim.omp_mult_vec(view, np.asfortranarray(np.full(S.shape_fb, 2.0)))
    assert out is view and np.shares_memory(out, eg)
    assert np.array_equal(eg[:, :, :, 3:], before)


def test_real_psatd_tables_are_cast_to_complex(ofim):
    """maxwell_push_wo_spchrg declares complex coefficients; the non-envelope driver passes float64 tables
    (solvers.py:244-279, maxwell_solvers.f90:67-68)"""
    S = setup("real_m3")
    assert S.PSATD_E.dtype == np.float64
    rng = np.random.default_rng(2)
    eg, j = crandn(rng, S.shape_fb + (6,)), crandn(rng, S.shape_fb + (3,))
    a = ofim.maxwell_push_wo_spchrg(eg.copy(order="F"), j, S.PSATD_E, S.PSATD_G)
    b = ofim.maxwell_push_wo_spchrg(eg.copy(order="F"), j, S.PSATD_E.astype(complex), S.PSATD_G.astype(complex))
    assert np.array_equal(a, b)
    # the cast is kept per table: a second call gives the same result, a rebuilt table (new array, e.g. after a change
    # of the time step) and an edit of the same array are both seen
    a2 = ofim.maxwell_push_wo_spchrg(eg.copy(order="F"), j, S.PSATD_E, S.PSATD_G)
    assert np.array_equal(a2, a)
    E2 = np.asfortranarray(S.PSATD_E * 0.5)
    c = ofim.maxwell_push_wo_spchrg(eg.copy(order="F"), j, E2, S.PSATD_G)
    d = ofim.maxwell_push_wo_spchrg(eg.copy(order="F"), j, E2.astype(complex), S.PSATD_G.astype(complex))
    assert np.array_equal(c, d) and not np.array_equal(c, a)
    E2 *= 3.0
    e = ofim.maxwell_push_wo_spchrg(eg.copy(order="F"), j, E2, S.PSATD_G)
    f = ofim.maxwell_push_wo_spchrg(eg.copy(order="F"), j, E2.astype(complex), S.PSATD_G.astype(complex))
    assert np.array_equal(e, f)


def test_both_backends_expose_the_same_api(ofim):
    import chimera_b200.fimera as g

    assert g.API_NAMES == ofim.API_NAMES
