"""CPU: the multi-GPU host logic on a world_size-2 gloo job and the pure
partitioning functions of chimera_b200/sharding.py."""
import copy
import os
import subprocess
import sys

import numpy as np
import pytest

from util import SETUPS, assert_close, match, plasma, seed_fields
from chimera_b200 import sharding
from chimera_b200.solver_setup import SolverSetup

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("n,world", [(10, 1), (10, 3), (7, 8), (98485632, 8)])
def test_particle_ranges_partition(n, world):
    r = [sharding.particle_range(n, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == n
    assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
    sizes = [b - a for a, b in r]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("nx,world", [(16, 1), (16, 2), (80, 2), (4096, 8), (48, 4)])
def test_kx_slabs_are_closed_under_the_mirror(nx, world):
    seen = np.zeros(nx, dtype=int)
    for rank in range(world):
        rows = sharding.kx_slab_rows(nx, rank, world)
        assert rows.size == nx // world and np.all(np.diff(rows) > 0)
        seen[rows] += 1
        # the local mirror formula reproduces the global one, fb_math.f90:35-36
        loc = sharding.local_mirror(rows.size, sharding.mirror_shift(rank, world))
        assert np.array_equal(rows[loc], (nx - rows) % nx)
    assert np.all(seen == 1)


def test_slab_needs_divisibility():
    assert not sharding.slab_supported(30, 4)
    with pytest.raises(ValueError):
        sharding.kx_slab_rows(30, 1, 4)


@pytest.fixture(scope="module")
def two_rank_run(tmp_path_factory):
    out = tmp_path_factory.mktemp("gloo")
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + os.getpid() % 2000), os.path.join(HERE, "dist_worker.py"), str(out)]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    return out


def test_particle_sharding_with_grid_allreduce_matches_one_rank(ofim, two_rank_run):
    from pic_ref import RefRun, RefSpecies

    S = SolverSetup(copy.deepcopy(SETUPS["real_m2"]))
    x, p, w = plasma(S, 2, 2, 41)
    xi, pi_, wi = plasma(S, 2, 2, 47)
    run = RefRun(ofim, S, [RefSpecies(x, p, w), RefSpecies(xi, 0 * pi_, -wi, charge=1.0, mass=1886.0, still=True)],
                 background=True)
    run.EG_fb[:] = seed_fields(S, 42)
    run.make_halfstep(px0=(0.0, 0.0))
    run.make_step()
    run.make_step()
    parts = [np.load(os.path.join(two_rank_run, "particles_%d.npz" % r)) for r in range(2)]
    for z in parts:  # the grids are replicated after the all-reduce
        assert_close(z["EG_fb"], run.EG_fb, 1e-12, "EG_fb")
        assert_close(z["J"], run.J, 1e-12, "J")
        assert_close(z["Rho"], run.Rho, 1e-12, "Rho")
    mom = np.concatenate([z["momenta"] for z in parts], axis=1)
    wts = np.concatenate([z["weights"] for z in parts])
    perm = match(run.sp[0].weights, wts)
    assert_close(mom[:, perm], run.sp[0].momenta, 1e-12, "momenta")


def test_kx_slab_spectral_update_matches_full_solve(ofim, two_rank_run):
    """3 Poisson iterations + PSATD advance + rot on two mirror-pair kx slabs (numpy restatement with the
    slab-local mirror) == the same on the full kx axis with the C++ oracle."""
    S = SolverSetup(copy.deepcopy(SETUPS["real_m2"]))
    a = S.Args
    rng = np.random.default_rng(5)
    shp = S.shape_fb
    J = np.asfortranarray(rng.standard_normal(shp + (3,)) + 1j * rng.standard_normal(shp + (3,)))
    g0 = np.asfortranarray(rng.standard_normal(shp + (3,)) + 1j * rng.standard_normal(shp + (3,)))
    g1 = np.asfortranarray(rng.standard_normal(shp + (3,)) + 1j * rng.standard_normal(shp + (3,)))
    EG = np.asfortranarray(rng.standard_normal(shp + (6,)) + 1j * rng.standard_normal(shp + (6,)))
    for _ in range(3):
        gd = ofim.fb_graddiv(J.copy(order="F"), *a["FBDiff"])
        J = ofim.poiss_corr(J, gd, g0, g1, a["dt_inv"], a["PoissFact"])
    EG = ofim.maxwell_push_with_spchrg(EG, J, g0, g1, S.PSATD_E, S.PSATD_G)
    B = ofim.omp_mult_vec(ofim.fb_rot(S.zeros_fb(3), EG[:, :, :, 3:], *a["FBDiff"]), a["PoissFact"])
    full = np.load(os.path.join(two_rank_run, "slab.npz"))["full"]
    assert_close(full[..., :6], EG, 1e-12, "EG_fb from slabs")
    assert_close(full[..., 6:], B, 1e-12, "B_fb from slabs")


@pytest.mark.parametrize("window,overlap,name", [(0, 1, "real_m2"), (0, 0, "real_m2"), (1, 1, "real_m2"), (1, 1, "env_m3")])
def test_engine_schedule_on_two_gloo_ranks(window, overlap, name):
    """the product's multi-rank host logic (Engine.make_halfstep / Engine.step in chimera_b200/engine.py: fused and
    unfused branches, re-binning cadence, rank-0 background rule, all-reduce of Rho behind fb_in_J, E and B halves of
    fields out all-gathered separately, per-step 'Staged' window; real solver with space charge and still ions, and
    the envelope solver without) on a world_size-2 gloo job, the CUDA library
    replaced by its CPU stand-in (tests/cpu_engine.py); every rank against the single-process reference sequence"""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(31500 + (os.getpid() + 7 * window + 3 * overlap + 17 * (name != "real_m2")) % 2000),
           os.path.join(HERE, "dist_engine_worker.py"), str(window), str(overlap), name]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "OK window" in r.stdout, r.stdout[-4000:]


def test_engine_lpa_window_on_two_gloo_ranks():
    """Engine.frame_act across ranks (damp_field over gathered kx slabs, added particles sharded, window cull at the
    species' upperR, background and density re-deposited and all-reduced) + Engine.step, replaying the LPA moving-window
    run recorded from the reference's own driver (tests/golden/real_m2_lpa.npz) on a world_size-2 gloo job"""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(33500 + os.getpid() % 2000), os.path.join(HERE, "dist_engine_worker.py"), "lpa"]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "OK lpa window" in r.stdout, r.stdout[-4000:]
