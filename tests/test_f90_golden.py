"""The pin: golden vectors produced by the reference's OWN Fortran (executed from /root/reference/f90 by
oracle/f90py.py, tools/gen_golden_f90.py) against the C++ oracle (CPU) and the CUDA library through its C ABI (GPU).

Every hot-path subroutine on seeded inputs (tests/f90_cases.py), plus make_halfstep + 2 make_step of the reference's
driver order.  Tolerances: 1e-13 for the oracle (different summation order only), 1e-12 for CUDA (north_star), integer
and index outputs exact; quantities that pass through the envelope carrier exp(+-i kx0 x) get util.carrier_tol."""
import os
import sys

import numpy as np
import pytest

import f90_cases
from f90_cases import fingerprint_error, flatten
from util import carrier_tol, setup

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = np.load(os.path.join(ROOT, "tests", "golden", "f90_kernels.npz"))
STEPS = np.load(os.path.join(ROOT, "tests", "golden", "f90_steps.npz"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def case_tol(cid, base):
    name = cid.split(":")[0]
    if name in f90_cases.SETUPS:
        t = carrier_tol(setup(name), base)
        # the envelope deposits sum terms that carry exp(-i kx0 x): a cancelling sum (test_gpu_engine.compare_state)
        return 20 * t if ("dep_" in cid and setup(name).env) else t
    return base


def check_backend(fim, base_tol):
    seen = 0
    worst = {}
    for cid, fn, args in f90_cases.cases(fim):
        res = getattr(fim, fn)(*[a.copy(order="F") if isinstance(a, np.ndarray) else a for a in args])
        for k, arr in enumerate(flatten(res)):
            err = fingerprint_error(arr, KERNELS["%s|%d" % (cid, k)])
            worst[cid] = max(worst.get(cid, 0.0), err)
            assert err <= case_tol(cid, base_tol), "%s output %d: error %.3e vs the reference's Fortran" % (cid, k, err)
            seen += 1
    assert seen == len(KERNELS.files), (seen, len(KERNELS.files))
    return worst


def test_oracle_matches_the_reference_fortran(ofim):
    worst = check_backend(ofim, 1e-13)
    assert len(worst) >= 110


@pytest.mark.gpu
def test_cuda_matches_the_reference_fortran(gfim):
    check_backend(gfim, 1e-12)


def run_steps(fim, base_tol):
    from gen_golden_f90 import build_step_run, step_cases, step_state

    for sid, name, kw in step_cases():
        run, px0 = build_step_run(fim, sid, name, **kw)
        tol = carrier_tol(run.S, base_tol)
        run.make_halfstep(px0=px0)
        for tag, nsteps in (("half", 0), ("step2", 2)):
            for _ in range(nsteps):
                run.make_step()
            for k, v in step_state(run).items():
                err = fingerprint_error(np.asarray(v), STEPS["%s|%s|%s" % (sid, tag, k)])
                t = 20 * tol if (run.env and k == "J") else tol
                assert err <= 5 * t, "%s after %s: %s differs from the reference's Fortran by %.3e" % (sid, tag, k, err)


def test_oracle_step_sequence_matches_the_reference_fortran(ofim):
    run_steps(ofim, 1e-13)


@pytest.mark.gpu
def test_cuda_step_sequence_matches_the_reference_fortran(gfim):
    run_steps(gfim, 1e-12)


# ---- the translator itself, on Fortran written for the purpose (no reference needed) --------------------------------
SNIPPET = """
subroutine probe(a, b, c, n, m, k)
implicit none
integer, intent(in) :: n, m
integer, intent(out) :: k
real (kind=8), intent(in) :: a(0:n, m)
real (kind=8), intent(inout) :: b(n+1)
complex(kind=8), intent(inout) :: c(-1:1)
integer :: i, j
real (kind=8) :: s, t(2)
complex(kind=8) :: ii=(0.0d0,1.0d0)
k = (n + 4) / 3            ! integer division truncates
s = 0.0d0
do i = 0, n
  do j = m, 1, -1          ! negative stride
    if (a(i,j) < 0.0d0) CYCLE
    s = s + a(i,j) * 2**j &
          - 1.5d0          ! continuation line
  enddo
enddo
b(1) = s
b(2:n+1) = a(n:1:-1, 1)    ! reversed section
t = 0.0d0
t(2) = SUM(ABS(a(0,:)))
if ((t(2) .ge. 0.0d0) .and. (.not. (n == 0))) then
  c(-1) = CONJG(c(1)) * ii
elseif (n == 0) then
  c(-1) = 0.0d0
else
  c(-1) = 1.0d0
endif
c(0) = DCMPLX(t(2), -t(2)) / 2
end subroutine
"""


def test_translator_on_a_known_snippet(tmp_path):
    from oracle.f90py import F90Module

    p = tmp_path / "probe.f90"
    p.write_text(SNIPPET)
    mod = F90Module([str(p)])
    rng = np.random.default_rng(3)
    n, m = 5, 3
    a = np.asfortranarray(rng.standard_normal((n + 1, m)))
    b = np.zeros(n + 1)
    c = np.asfortranarray(rng.standard_normal(3) + 1j * rng.standard_normal(3))
    c0 = c.copy()
    out = mod.probe(a, b, c, n, m, 0)
    assert out["k"] == 3
    s = 0.0
    for i in range(n + 1):
        for j in range(m, 0, -1):
            if a[i, j - 1] < 0:
                continue
            s = s + a[i, j - 1] * 2 ** j - 1.5
    assert b[0] == s
    assert np.array_equal(b[1:], a[n:0:-1, 0])
    t2 = np.abs(a[0]).sum()
    assert c[0] == np.conj(c0[2]) * 1j
    assert np.isclose(c[1], complex(t2, -t2) / 2, rtol=1e-15)
    assert c[2] == c0[2]


def test_translator_refuses_out_of_bounds(tmp_path):
    """the reference has no bounds checks; the executor does, so fixtures never contain undefined behaviour"""
    from oracle.f90py import F90Module

    p = tmp_path / "oob.f90"
    p.write_text("subroutine oob(a, n)\nimplicit none\ninteger, intent(in) :: n\nreal (kind=8), intent(inout) :: a(n)\na(n+1) = 1.0d0\nend subroutine\n")
    mod = F90Module([str(p)])
    with pytest.raises(IndexError):
        mod.oob(np.zeros(4), 4)


@pytest.mark.skipif(not os.path.isdir("/root/reference/f90"), reason="needs the reference checkout (build container)")
def test_fixtures_are_what_the_reference_fortran_gives_today():
    """regenerate a third of the kernel fixtures from /root/reference and compare bit for bit"""
    from oracle import fimera_f90

    F = fimera_f90.load()
    assert not {k: v for k, v in F._f90.warnings.items() if v}
    for n, (cid, fn, args) in enumerate(f90_cases.cases(F)):
        if n % 3:
            continue
        res = getattr(F, fn)(*[a.copy(order="F") if isinstance(a, np.ndarray) else a for a in args])
        for k, arr in enumerate(flatten(res)):
            assert np.array_equal(f90_cases.fingerprint(arr), KERNELS["%s|%d" % (cid, k)]), cid


# ---- the WHOLE reference stack: its unmodified Python driver on top of its own Fortran ------------------------------
@pytest.mark.skipif(not os.path.isdir("/root/reference/f90"), reason="needs the reference checkout (build container)")
@pytest.mark.parametrize("name", ["real_m2", "real_m3", "env_m1", "env_m3", "static_m2", "env_m1_win"])
def test_reference_driver_on_its_own_fortran_reproduces_the_fixtures(name):
    """tests/golden/<name>.npz were recorded by running the reference's ChimeraRun / Solver / Specie on the C++ oracle
    (tools/gen_golden.py).  Here the same unmodified driver runs on the reference's OWN Fortran (oracle/f90py.py): every
    recorded state (after make_halfstep and after 4 make_step: grids, spectral fields, particles, and for the windowed FEL
    case the reference's Diagnostics) must come out the same, so those fixtures are outputs of the reference itself."""
    import gen_golden
    import ref_driver
    from oracle import fimera_f90

    F = fimera_f90.load()
    R = ref_driver.install(F)
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        out = gen_golden.generate(name, gen_golden.CASES[name], R, F)
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    S = setup(gen_golden.CASES[name]["setup"])
    tol = carrier_tol(S, 1e-12)
    checked = 0
    for k in gold.files:
        if not k.startswith(("h_", "s_", "d_", "tab_")) or gold[k].dtype.kind not in "fc":
            continue
        a, b = np.asarray(out[k]), gold[k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        den = np.linalg.norm(b.ravel())
        err = np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0)
        t = 20 * tol if (S.env and k.endswith("_J")) else tol
        assert err <= t, "%s/%s: the reference on its own Fortran differs from the fixture by %.3e" % (name, k, err)
        checked += 1
    assert checked >= 8, checked
