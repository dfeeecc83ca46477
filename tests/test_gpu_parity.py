"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the
same inputs.  Floating point: solution within 1e-10 relative max-abs of the
oracle's, iteration count within +-2 (the north star's tolerance)."""

from __future__ import annotations

import numpy as np
import pytest

from aphros_b200 import Conf, Mesh, SolverConjugateCuda, SolverJacobiCuda, capi
from cases import (case_density, case_periodic_const, case_tlinear, iteration_budget,
                   iterations_ok, random_guess, rel_max_abs, remove_mean)

pytestmark = pytest.mark.gpu

X_TOL = 1e-10   # relative max-abs error of the solution (north star)
ITER_TOL = 2    # iterations to tolerance


def oracle_solve(case, **kw):
    from oracle import cpu
    return cpu.solve(case["system"], kw.pop("x0", None), periodic=case["periodic"], **kw)


def gpu_solve(case, conf, x0=None, maxnorm=False, flags=0):
    shape = case["system"].shape[:3]
    solver = SolverConjugateCuda(conf, {"residual_max": maxnorm},
                                 Mesh(shape=shape, periodic=case["periodic"]), flags)
    x = np.full(shape, np.nan)
    info = solver.Solve(case["system"], x0, x)
    hist = solver.History(info.iter)
    solver.close()
    return x, info, hist


def compare_solutions(case, x, xo):
    # singular (pure Neumann / periodic) systems fix the solution up to a constant:
    # the iteration never changes the mean, so both start and stay on the same one
    return rel_max_abs(x, xo)


VARIANTS = [pytest.param(0, id="tma-stream+sym+graph"),
            pytest.param(capi.APHCG_NO_SYM, id="tma+7coef+graph"),
            pytest.param(capi.APHCG_NO_TMA, id="plain+graph"),
            pytest.param(capi.APHCG_NO_STREAM, id="tma-ldg+sym+graph"),
            pytest.param(capi.APHCG_NO_TMA | capi.APHCG_NO_GRAPH, id="plain+nograph"),
            pytest.param(capi.APHCG_NO_GRAPH, id="tma+nograph")]


@pytest.mark.parametrize("flags", VARIANTS)
def test_tlinear_32_converged(gpu, flags):
    """The reference's own unit test system (src/test/linear/main.cpp), its command
    line: --tol 1e-5 --maxiter 1000 (src/test/linear/test:12-16)."""
    case = case_tlinear(32)
    conf = Conf(tol=1e-5, miniter=0, maxiter=1000)
    x, info, hist = gpu_solve(case, conf, flags=flags)
    xo, it_o, res_o, hist_o = oracle_solve(case, tol=conf.tol, miniter=0, maxiter=conf.maxiter)
    assert abs(info.iter - it_o) <= ITER_TOL
    assert info.residual < conf.tol
    # the reference's own acceptance check: error vs the exact solution
    d = remove_mean(x - case["exact"])
    do = remove_mean(xo - case["exact"])
    assert np.abs(d).max() < 2 * np.abs(do).max() + 1e-12
    n = min(len(hist), len(hist_o), 50)
    np.testing.assert_allclose(hist[:n], hist_o[:n], rtol=1e-9)


PARITY_CASES = {
    "tlinear48": lambda: case_tlinear(48),
    "const40": lambda: case_periodic_const(40),
    "density32_10to1": lambda: case_density(32, rho_in=0.1),
    "density_perz": lambda: case_density(32, periodic=(False, False, True), rho_in=0.1),
    "flat": lambda: case_tlinear(None, shape=(4, 36, 130)),
}


def rhs_norm_of(case):
    """sqrt(sum e7^2 / V): the residual of the zero guess, in the reference's norm"""
    shape = case["system"].shape[:3]
    return float(np.sqrt((case["system"][..., 7] ** 2).sum() / __import__('aphros_b200').systems.cell_volume(shape)))


@pytest.mark.parametrize("flags", VARIANTS[:4])
@pytest.mark.parametrize("name", sorted(PARITY_CASES))
def test_iterations_to_tolerance(gpu, name, flags):
    """iteration count to a 1e-8 relative residual (the north star's setting,
    expressed as the absolute tol the reference takes) within +-2 -- plus, on the
    variable-density systems, the reference's own summation-order spread (its count
    moves with the block size there; 0 on the well-conditioned systems)"""
    case = PARITY_CASES[name]()
    conf = Conf(tol=1e-8 * rhs_norm_of(case), miniter=0, maxiter=5000)
    x, info, hist = gpu_solve(case, conf, flags=flags)
    xo, it_o, res_o, hist_o = oracle_solve(case, tol=conf.tol, miniter=0, maxiter=conf.maxiter)
    assert it_o < conf.maxiter, "oracle did not converge: bad test case"
    if name.startswith("density"):
        _, counts = iteration_budget(case["system"], None, case["periodic"], conf.tol,
                                     conf.maxiter, blocks=(4, 8, 16, 32))
        assert iterations_ok(info.iter, counts + [it_o]), (info.iter, it_o, counts)
    else:
        assert abs(info.iter - it_o) <= ITER_TOL, (info.iter, it_o)
    assert info.residual < conf.tol


@pytest.mark.parametrize("flags", VARIANTS[:4])
@pytest.mark.parametrize("name", sorted(PARITY_CASES))
def test_solution_parity(gpu, name, flags):
    """solution within 1e-10 relative max-abs of the oracle's once both are
    converged (tol = 1e-12 relative).  So close to the rounding floor the
    iteration count depends on the summation order -- the reference itself moves
    by +-4 iterations between block sizes -- so here it is only
    bounded by the oracle's own block-8-vs-32 spread plus the +-2 budget."""
    case = PARITY_CASES[name]()
    conf = Conf(tol=1e-12 * rhs_norm_of(case), miniter=0, maxiter=5000)
    x, info, hist = gpu_solve(case, conf, flags=flags)
    xo, it_o, res_o, hist_o = oracle_solve(case, tol=conf.tol, miniter=0, maxiter=conf.maxiter)
    assert it_o < conf.maxiter, "oracle did not converge: bad test case"
    assert compare_solutions(case, x, xo) <= X_TOL
    shape = case["system"].shape[:3]
    blk = tuple(max(1, min(8, n)) for n in shape[::-1])
    _, it_b, _, _ = oracle_solve(case, tol=conf.tol, miniter=0, maxiter=conf.maxiter, block=blk)
    assert abs(info.iter - it_o) <= ITER_TOL + abs(it_b - it_o) + it_o // 100, (info.iter, it_o, it_b)


def test_fixed_iterations_history(gpu):
    """tol=0, maxiter=100 (the reference's run_bench setting) -> exactly 101
    iterations (src/linear/linear.ipp:110-113) and the same residual history."""
    case = case_tlinear(64)
    conf = Conf(tol=0.0, miniter=0, maxiter=100)
    x, info, hist = gpu_solve(case, conf)
    xo, it_o, res_o, hist_o = oracle_solve(case, tol=0.0, miniter=0, maxiter=100)
    assert info.iter == it_o == 101
    # known answer regenerated from the reference: res=4.199885e-01
    assert abs(res_o - 4.199885e-01) < 1e-6
    np.testing.assert_allclose(hist, hist_o, rtol=1e-7)
    assert rel_max_abs(x, xo) < 1e-8


def test_initial_guess_and_alias(gpu):
    case = case_tlinear(32)
    x0 = random_guess(case["system"].shape[:3])
    conf = Conf(tol=1e-7, miniter=0, maxiter=2000)
    xo, it_o, _, _ = oracle_solve(case, x0=x0, tol=conf.tol, miniter=0, maxiter=conf.maxiter)
    shape = x0.shape
    solver = SolverConjugateCuda(conf, {}, Mesh(shape=shape, periodic=case["periodic"]))
    x = x0.copy()
    info = solver.Solve(case["system"], x, x)  # fc_init == &fc_sol (linear.h:40)
    assert abs(info.iter - it_o) <= ITER_TOL
    assert rel_max_abs(x, xo) <= X_TOL
    # SetConf between solves (hydro.ipp:2351-2353), zero guess
    solver.SetConf(Conf(tol=1e-3, miniter=0, maxiter=2000))
    x2 = np.zeros(shape)
    info2 = solver.Solve(case["system"], None, x2)
    xo2, it_o2, _, _ = oracle_solve(case, tol=1e-3, miniter=0, maxiter=2000)
    assert abs(info2.iter - it_o2) <= ITER_TOL
    assert rel_max_abs(x2, xo2) < 1e-9
    solver.close()


def test_miniter_maxiter_rules(gpu):
    case = case_tlinear(16)
    for conf in [Conf(tol=1e30, miniter=7, maxiter=100),   # converged at once, miniter wins
                 Conf(tol=0.0, miniter=0, maxiter=5),      # maxiter+1 iterations
                 Conf(tol=0.0, miniter=9, maxiter=3),      # miniter > maxiter
                 Conf(tol=1e30, miniter=0, maxiter=0)]:    # one iteration at least
        x, info, _ = gpu_solve(case, conf)
        xo, it_o, res_o, _ = oracle_solve(case, tol=conf.tol, miniter=conf.miniter,
                                          maxiter=conf.maxiter)
        assert info.iter == it_o, (conf, info.iter, it_o)
        assert rel_max_abs(x, xo) < 1e-10
        assert abs(info.residual - res_o) <= 1e-9 * abs(res_o)


def test_maxnorm(gpu):
    case = case_tlinear(32)
    conf = Conf(tol=1e-2, miniter=0, maxiter=1000)
    x, info, hist = gpu_solve(case, conf, maxnorm=True)
    xo, it_o, res_o, hist_o = oracle_solve(case, tol=conf.tol, miniter=0, maxiter=1000, maxnorm=True)
    assert abs(info.iter - it_o) <= ITER_TOL
    n = min(len(hist), len(hist_o), 30)
    np.testing.assert_allclose(hist[:n], hist_o[:n], rtol=1e-8)


@pytest.mark.parametrize("shape", [(1, 1, 2), (3, 5, 7), (5, 9, 130), (2, 8, 128), (9, 16, 258),
                                   (33, 8, 64), (1, 40, 40), (2, 8, 1024), (3, 10, 1100),
                                   (2, 1030, 64), (70, 8, 640)])
def test_ragged_shapes(gpu, shape):
    """odd nx (scalar path), partial tiles, single plane (2-D), tiny meshes"""
    s, exact = __import__("aphros_b200").systems.tlinear_system(None, shape=shape)
    case = dict(system=s, periodic=(True, True, True))
    conf = Conf(tol=0.0, miniter=0, maxiter=12)
    x, info, hist = gpu_solve(case, conf)
    xo, it_o, res_o, hist_o = oracle_solve(case, tol=0.0, miniter=0, maxiter=12)
    assert info.iter == it_o
    # compare the history down to 1e-9 of the initial residual (below that both
    # sit on the rounding floor, where digits are noise)
    from cases import initial_residual
    res0 = initial_residual(s, None, (True, True, True))
    live = hist_o > 1e-9 * res0
    k = int(np.argmin(live)) if not live.all() else len(hist_o)
    np.testing.assert_allclose(hist[:k], hist_o[:k], rtol=1e-7)


def test_apply_operator(gpu):
    """stage "iter" operator alone (linear.ipp:65-72): differences only from FMA"""
    from oracle import cpu
    for case in [case_tlinear(24), case_density(24, periodic=(False, True, False)),
                 case_tlinear(None, shape=(5, 9, 130))]:
        shape = case["system"].shape[:3]
        v = random_guess(shape, seed=11)
        solver = SolverConjugateCuda(Conf(), {}, Mesh(shape=shape, periodic=case["periodic"]))
        solver.UploadSystem(case["system"])
        out = solver.Apply(v)
        ref = cpu.apply(case["system"], v, periodic=case["periodic"])
        scale = cpu.apply(np.abs(case["system"]), np.abs(v), periodic=case["periodic"])
        assert (np.abs(out - ref) <= 8 * np.finfo(float).eps * scale + 1e-300).all()
        solver.close()


def test_strided_fields(gpu):
    """fields laid out like the reference's FieldCell with hl=2 halos and one padding
    cell per direction (src/geom/mesh.ipp:60-113)"""
    case = case_tlinear(16)
    n, hl = 16, 2
    full = n + 2 * hl + 1
    sys_full = np.full((full, full, full, 8), np.nan)
    sys_full[hl:hl + n, hl:hl + n, hl:hl + n] = case["system"]
    x_full = np.full((full, full, full), 777.0)
    conf = Conf(tol=1e-8, miniter=0, maxiter=1000)
    solver = SolverConjugateCuda(conf, {}, Mesh(shape=(n, n, n), periodic=case["periodic"]))
    xv = x_full[hl:hl + n, hl:hl + n, hl:hl + n]
    xv[...] = 0
    info = solver.Solve(sys_full[hl:hl + n, hl:hl + n, hl:hl + n], xv, xv)
    solver.close()
    xo, it_o, _, _ = oracle_solve(case, tol=conf.tol, miniter=0, maxiter=1000)
    assert abs(info.iter - it_o) <= ITER_TOL
    assert rel_max_abs(xv, xo) <= X_TOL
    mask = np.ones_like(x_full, dtype=bool)
    mask[hl:hl + n, hl:hl + n, hl:hl + n] = False
    assert (x_full[mask] == 777.0).all(), "cells outside the inner block were touched"


def test_deterministic(gpu):
    case = case_density(32, rho_in=0.01)
    conf = Conf(tol=0.0, miniter=0, maxiter=60)
    x1, i1, h1 = gpu_solve(case, conf)
    x2, i2, h2 = gpu_solve(case, conf)
    assert np.array_equal(x1, x2) and np.array_equal(h1, h2)


def test_jacobi(gpu):
    """SolverJacobi twin (linear.ipp:152-237); known answer of the reference test:
    32^3, tol 1e-5 -> iter=613 res=9.981133e-06"""
    from oracle import cpu
    case = case_tlinear(32)
    conf = Conf(tol=1e-5, miniter=0, maxiter=1000)
    solver = SolverJacobiCuda(conf, {}, Mesh(shape=(32, 32, 32), periodic=case["periodic"]))
    x = np.zeros((32, 32, 32))
    info = solver.Solve(case["system"], None, x)
    solver.close()
    xo, it_o, res_o, _ = cpu.solve(case["system"], tol=1e-5, miniter=0, maxiter=1000, method="jacobi")
    assert it_o == 613 and abs(res_o - 9.981133e-06) < 1e-11
    assert abs(info.iter - it_o) <= ITER_TOL
    assert rel_max_abs(x, xo) < 1e-9


def test_device_assembly_matches_host_generator(gpu):
    from aphros_b200 import systems
    n = 48
    sph = systems.random_spheres(16, 5)
    ref, _ = systems.density_poisson_system(n, nspheres=16, seed=5)
    solver = SolverConjugateCuda(Conf(), {}, Mesh(shape=(n, n, n), periodic=(False, False, False)))
    solver.AssembleSpheres(sph)
    got = solver.DownloadSystem()
    solver.close()
    assert np.array_equal(got[..., :7], ref[..., :7])
    # right-hand side: device sin/cos differ from libm by a few ulp of the fluxes,
    # and e7 is a difference of neighbouring fluxes (cancellation ~ 1/h)
    scale = np.abs(ref[..., 7]).max()
    assert np.abs(got[..., 7] - ref[..., 7]).max() <= 1e-12 * scale



def test_nonsymmetric_storage_falls_back(gpu):
    """a system whose off-diagonals are not bitwise symmetric must take the 7-stream
    path and still match the oracle (CG itself may not converge on it: compare a fixed
    number of iterations)"""
    case = case_tlinear(24)
    s = case["system"].copy()
    s[3, 4, 5, 2] *= 1.0 + 2.0 ** -40      # x+ of one cell, no longer equal to its mirror
    s[7, 0, 9, 4] *= 1.0 - 2.0 ** -41      # y+
    s[11, 2, 2, 6] *= 1.0 + 2.0 ** -39     # z+
    case = dict(system=s, periodic=case["periodic"])
    conf = Conf(tol=0.0, miniter=0, maxiter=40)
    x, info, hist = gpu_solve(case, conf)
    xo, it_o, res_o, hist_o = oracle_solve(case, tol=0.0, miniter=0, maxiter=40)
    assert info.iter == it_o == 41
    np.testing.assert_allclose(hist, hist_o, rtol=1e-9)
    assert rel_max_abs(x, xo) < 1e-10
    # and the perturbation is visible at all: the symmetric path would differ
    x_sym, _, hist_sym = gpu_solve(dict(system=case_tlinear(24)["system"], periodic=case["periodic"]), conf)
    assert not np.array_equal(hist_sym, hist)


def test_tlinear_cli(gpu, tmp_path, capsys):
    """the t.linear command line of the reference test (src/test/linear/test:12-16:
    --tol 1e-5 --maxiter 1000 --verbose --solver <case>) and its acceptance check
    max_diff_exact < 2*ref + 1e-12 with ref/conjugate/out: max_diff_exact=6.926950e-03"""
    import re
    from aphros_b200 import tlinear
    sysfile = str(tmp_path / "sys.raw")
    tlinear.main(["--tol", "1e-5", "--maxiter", "1000", "--verbose", "--solver", "conjugate_cuda",
                  "--system_out", sysfile])
    out = capsys.readouterr().out
    got = dict(re.findall(r"(\w+)=(\S+)", out))
    assert float(got["max_diff_exact"]) < 2 * 6.926950e-03 + 1e-12
    assert abs(int(got["iter"]) - 122) <= 2            # current reference code: iter=122
    assert abs(float(got["max_diff_exact"]) - 3.10655e-07) < 1e-10
    # replay the captured system
    tlinear.main(["--tol", "1e-5", "--maxiter", "1000", "--verbose", "--solver", "conjugate_cuda",
                  "--system_in", sysfile])
    got2 = dict(re.findall(r"(\w+)=(\S+)", capsys.readouterr().out))
    assert got2["iter"] == got["iter"] and got2["residual"] == got["residual"]
    tlinear.main(["--tol", "1e-5", "--maxiter", "1000", "--verbose", "--solver", "jacobi_cuda"])
    got3 = dict(re.findall(r"(\w+)=(\S+)", capsys.readouterr().out))
    assert abs(int(got3["iter"]) - 613) <= 2


# ---- opt-in Jacobi-preconditioned mode (APHCG_JACOBI_PRECOND) -----------------------
# Not a reference algorithm (SolverConjugate is unpreconditioned): the
# checker is our own restatement oracle/cg_oracle.c:cg_oracle_pconjugate ("parity
# unpinned"), plus the property that it coincides with plain CG when diag(A) is constant.

def gpu_solve_precond(case, conf, x0=None, flags=0):
    shape = case["system"].shape[:3]
    solver = SolverConjugateCuda(conf, {"jacobi_precond": True},
                                 Mesh(shape=shape, periodic=case["periodic"]), flags)
    x = np.full(shape, np.nan)
    info = solver.Solve(case["system"], x0, x)
    hist = solver.History(info.iter)
    solver.close()
    return x, info, hist


def test_precond_equals_plain_cg_for_constant_diagonal(gpu):
    case = case_periodic_const(32)
    conf = Conf(tol=1e-10 * rhs_norm_of(case), miniter=0, maxiter=2000)
    xp, ip, hp = gpu_solve_precond(case, conf)
    x, i, h = gpu_solve(case, conf)
    assert ip.iter == i.iter
    assert rel_max_abs(xp, x) <= 1e-10


@pytest.mark.parametrize("flags", [0, capi.APHCG_NO_TMA], ids=["tma", "plain"])
def test_precond_matches_its_oracle(gpu, flags):
    from oracle import cpu
    for case, x0 in [(case_density(32, rho_in=1e-3), None),
                     (case_tlinear(32, rho_in=100.0), random_guess((32, 32, 32)))]:
        from cases import initial_residual
        tol = 1e-9 * initial_residual(case["system"], x0, case["periodic"])
        conf = Conf(tol=tol, miniter=0, maxiter=5000)
        x, info, hist = gpu_solve_precond(case, conf, x0=x0, flags=flags)
        xo, it_o, res_o, hist_o = cpu.solve(case["system"], x0, periodic=case["periodic"], tol=tol,
                                            miniter=0, maxiter=5000, method="pconjugate")
        xc, it_c, _, _ = cpu.solve(case["system"], x0, periodic=case["periodic"], tol=tol,
                                   miniter=0, maxiter=20000)
        assert it_o < 5000
        assert abs(info.iter - it_o) <= 2 + it_o // 50, (info.iter, it_o)
        assert info.iter < it_c, "the preconditioner should pay off on a variable-density system"
        n = min(len(hist), len(hist_o), 40)
        np.testing.assert_allclose(hist[:n], hist_o[:n], rtol=1e-6)
        assert rel_max_abs(remove_mean(x), remove_mean(xo)) <= 1e-6


def test_device_assembly_from_density_and_fluxes(gpu):
    """aphcg_assemble_projection: rows assembled on the device from a
    cell density and face fluxes are bit for bit the host statement's
    (systems.projection_rows, itself pinned to the reference's assembler in test_oracle.py);
    periodic in z and walls in x, y to exercise wrap and clamp"""
    from aphros_b200 import systems
    shape = (12, 10, 16)
    nz, ny, nx = shape
    per = (False, False, True)
    rng = np.random.default_rng(3)
    rho = np.exp(rng.standard_normal(shape))
    h, dt = 1.0 / max(shape), 1e-3
    vx = rng.standard_normal((nz, ny, nx + 1))
    vy = rng.standard_normal((nz, ny + 1, nx))
    vz = rng.standard_normal((nz + 1, ny, nx))
    vz[-1] = vz[0]                       # periodic in z: same face
    src = rng.standard_normal(shape)
    vol = systems.cell_volume(shape)
    ref = systems.projection_rows(rho, vx, vy, vz, src, dt=dt, periodic=per, h=h, volume=vol)
    rho_ext = np.concatenate([rho[-1:], rho, rho[:1]])   # z ghost planes (periodic images)
    solver = SolverConjugateCuda(Conf(), {}, Mesh(shape=shape, periodic=per))
    solver.AssembleProjection(rho_ext, vx, vy, vz, source=src, dt=dt, h=h)
    got = solver.DownloadSystem()
    solver.close()
    assert np.array_equal(got[..., :7], ref[..., :7])
    assert np.array_equal(got[..., 7], ref[..., 7])


def test_error_paths(gpu):
    """call-order and argument errors come back as codes + messages, never as crashes
    (the adapter turns them into fassert, src/util/logger.h:44-60)"""
    import ctypes
    L = capi.lib()
    solver = SolverConjugateCuda(Conf(tol=1e-6, miniter=0, maxiter=10), {}, Mesh(shape=(8, 8, 8)))
    info = capi.Info()
    c = capi.Conf(1e-6, 0, 10)
    assert L.aphcg_run(solver._h, ctypes.byref(c), ctypes.byref(info)) == -4      # no system yet
    assert b"no system" in L.aphcg_last_error()
    with pytest.raises(ValueError):
        solver.Solve(np.zeros((8, 8, 8, 7)), None, np.zeros((8, 8, 8)))           # wrong row width
    with pytest.raises(ValueError):
        solver.Solve(np.zeros((8, 8, 8, 8), dtype=np.float32), None, np.zeros((8, 8, 8)))
    bad = capi.Layout(0, 4, 64)                                                    # stride_y < nx
    assert L.aphcg_upload_guess(solver._h, capi.ptr(np.zeros(512)), ctypes.byref(bad)) == -1
    c_bad = capi.Conf(0.0, 0, -1)
    s, _ = case_tlinear(8)["system"], None
    solver.UploadSystem(s)
    assert L.aphcg_run(solver._h, ctypes.byref(c_bad), ctypes.byref(info)) == -1   # negative maxiter
    # and the handle is still usable afterwards
    x = np.zeros((8, 8, 8))
    assert solver.Solve(s, None, x).iter >= 1
    solver.close()
    solver.close()  # idempotent


@pytest.mark.parametrize("precond", [False, True])
def test_initial_residual_norm_is_reported(gpu, precond):
    """aphcg_info.residual0 = sqrt(sum r0^2 / V), r0 = -(A x0 + e7) (linear.ipp:48-56): what a
    caller needs to turn a relative tolerance into the reference's absolute one"""
    from cases import initial_residual
    from aphros_b200 import SolverConjugateCudaGroup
    case = case_density(24, rho_in=0.1)
    shape = case["system"].shape[:3]
    x0 = random_guess(shape)
    want = initial_residual(case["system"], x0, case["periodic"])
    m = Mesh(shape=shape, periodic=case["periodic"])
    conf = Conf(tol=0.0, miniter=0, maxiter=3)
    for make in (lambda: SolverConjugateCuda(conf, {"jacobi_precond": precond}, m),
                 lambda: SolverConjugateCudaGroup(conf, {"jacobi_precond": precond}, m, [0, 0, 0])):
        solver = make()
        info = solver.Solve(case["system"], x0, np.zeros(shape))
        solver.close()
        assert abs(info.residual0 - want) <= 1e-12 * want
    want0 = rhs_norm_of(case)
    solver = SolverConjugateCuda(conf, {"jacobi_precond": precond}, m)
    info = solver.Solve(case["system"], None, np.zeros(shape))
    solver.close()
    assert abs(info.residual0 - want0) <= 1e-12 * want0


def test_batched_x_update_is_bitwise_the_per_iteration_update(gpu, monkeypatch):
    """The TMA kernel applies the deferred `x += alpha p` (linear.ipp:88) two iterations at a
    time, in iteration order with the same FMAs; APHCG_XBATCH=0 applies one per iteration.
    Same bits for odd and even iteration counts, with and without an initial guess."""
    case = case_density(32, rho_in=0.01)
    shape = case["system"].shape[:3]
    x0 = random_guess(shape)
    for maxiter in (0, 1, 2, 37, 60):
        out = []
        for xb in ("1", "0"):
            monkeypatch.setenv("APHCG_XBATCH", xb)
            for guess in (None, x0):
                x, info, _ = gpu_solve(case, Conf(tol=0.0, miniter=0, maxiter=maxiter), x0=guess)
                assert info.iter == maxiter + 1
                out.append(x)
        assert np.array_equal(out[0], out[2]) and np.array_equal(out[1], out[3]), maxiter
