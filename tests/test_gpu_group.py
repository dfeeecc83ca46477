"""In-process slab group (aphcg_group_*, include/aphcg.h): one process drives several
z-slabs -- the mode the aphros adapter uses for `cuda_devices > 1`.  The slab loop is the
multi-GPU loop (residual ghost planes and the scalars through peer memory, written by the
kernels), so giving the same device ordinal several times exercises the whole slab path on
a single-GPU machine; with >= 2 GPUs the same tests also run with one slab per GPU.
Everything is compared with the single-domain CPU oracle."""

from __future__ import annotations

import numpy as np
import pytest

from aphros_b200 import (Conf, Mesh, SolverConjugateCuda, SolverConjugateCudaGroup,
                         SolverJacobiCudaGroup, capi)
from aphros_b200.solver import ModuleLinear
from cases import (case_density, case_tlinear, initial_residual, iteration_budget, iterations_ok,
                   random_guess, rel_max_abs, solution_budget)

pytestmark = pytest.mark.gpu

X_TOL = 1e-10
ITER_TOL = 2


def device_lists():
    out = [pytest.param([0, 0], id="2slabs-1gpu"), pytest.param([0, 0, 0], id="3slabs-1gpu")]
    n = capi.device_count()
    if n >= 2:
        out.append(pytest.param([0, 1], id="2gpus"))
    if n >= 4:
        out.append(pytest.param([0, 1, 2, 3], id="4gpus"))
    if n >= 8:
        out.append(pytest.param(list(range(8)), id="8gpus"))
    return out


def oracle_solve(case, x0=None, **kw):
    from oracle import cpu
    return cpu.solve(case["system"], x0, periodic=case["periodic"], **kw)


@pytest.mark.parametrize("devices", device_lists())
def test_group_matches_oracle_tlinear(gpu, devices):
    """the reference's unit-test system, periodic in z too (the slab ring wraps)"""
    case = case_tlinear(32)
    shape = case["system"].shape[:3]
    x0 = random_guess(shape) * 1e-3
    tol = 1e-10 * initial_residual(case["system"], x0, case["periodic"])
    conf = Conf(tol=tol, miniter=0, maxiter=3000)
    solver = SolverConjugateCudaGroup(conf, {}, Mesh(shape=shape, periodic=case["periodic"]),
                                      devices)
    assert sum(n for _, n in solver.Slabs()) == shape[0]
    assert capi.lib().aphcg_group_size(solver._g) == len(devices)
    x = x0.copy()
    info = solver.Solve(case["system"], x, x)  # fc_init == &fc_sol (linear.h:40)
    xo, it_o, res_o, hist_o = oracle_solve(case, x0, tol=tol, miniter=0, maxiter=3000)
    assert abs(info.iter - it_o) <= ITER_TOL
    assert rel_max_abs(x, xo) <= X_TOL
    hist = solver.History(info.iter)
    n = min(len(hist), len(hist_o), 40)
    np.testing.assert_allclose(hist[:n], hist_o[:n], rtol=1e-9)
    # second solve on the same group: SetConf, zero guess, fixed iteration count
    solver.SetConf(Conf(tol=0.0, miniter=0, maxiter=30))
    x2 = np.full(shape, np.nan)
    info2 = solver.Solve(case["system"], None, x2)
    xo2, it_o2, res_o2, _ = oracle_solve(case, tol=0.0, miniter=0, maxiter=30)
    assert info2.iter == it_o2 == 31
    assert abs(info2.residual - res_o2) <= 1e-7 * res_o2
    assert rel_max_abs(x2, xo2) < 1e-8
    solver.close()


@pytest.mark.parametrize("devices", device_lists())
def test_group_variable_density_walls(gpu, devices):
    """Neumann walls, 20:1 density jump, uneven slabs (nz = 50 over 3 slabs -> 17/17/16)"""
    shape = (50, 24, 64)
    case = case_density(None, nspheres=6, seed=4, rho_in=0.05, shape=shape)
    x0 = random_guess(shape, seed=1) * 1e-3
    tol = 1e-10 * initial_residual(case["system"], x0, case["periodic"])
    conf = Conf(tol=tol, miniter=0, maxiter=4000)
    solver = SolverConjugateCudaGroup(conf, {}, Mesh(shape=shape, periodic=case["periodic"]),
                                      devices)
    x = np.zeros(shape)
    info = solver.Solve(case["system"], x0, x)
    solver.close()
    xo, it_o, _, _ = oracle_solve(case, x0, tol=tol, miniter=0, maxiter=4000)
    # ill-conditioned: iteration count and solution are defined up to the reference's own
    # summation-order spread (tests/cases.py)
    _, counts = iteration_budget(case["system"], x0, case["periodic"], tol, 4000, blocks=(4, 8, 16))
    xbudget, _ = solution_budget(case["system"], x0, case["periodic"], tol, 4000, blocks=(4, 8, 16))
    assert iterations_ok(info.iter, counts + [it_o]), (info.iter, counts, it_o)
    assert rel_max_abs(x, xo) <= xbudget


def test_group_equals_single_handle_when_one_slab(gpu):
    """a group of one slab is the single-GPU solver, bit for bit"""
    case = case_density(32, rho_in=0.1)
    shape = case["system"].shape[:3]
    conf = Conf(tol=0.0, miniter=0, maxiter=50)
    m = Mesh(shape=shape, periodic=case["periodic"])
    a = SolverConjugateCuda(conf, {}, m)
    xa = np.zeros(shape)
    ia = a.Solve(case["system"], None, xa)
    a.close()
    b = SolverConjugateCudaGroup(conf, {}, m, [0])
    xb = np.zeros(shape)
    ib = b.Solve(case["system"], None, xb)
    b.close()
    assert ia.iter == ib.iter and ia.residual == ib.residual
    assert np.array_equal(xa, xb)


def test_group_is_deterministic_and_slab_count_changes_only_rounding(gpu):
    case = case_tlinear(32)
    shape = case["system"].shape[:3]
    conf = Conf(tol=0.0, miniter=0, maxiter=40)
    m = Mesh(shape=shape, periodic=case["periodic"])
    res = []
    for devices in ([0, 0], [0, 0], [0, 0, 0, 0]):
        s = SolverConjugateCudaGroup(conf, {}, m, devices)
        x = np.zeros(shape)
        info = s.Solve(case["system"], None, x)
        s.close()
        res.append((x, info))
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1].residual == res[1][1].residual
    assert rel_max_abs(res[2][0], res[0][0]) < 1e-9


def test_group_strided_rank_wide_fields(gpu):
    """rank-wide arrays laid out like the reference's FieldCell (hl = 2 halos + one padding
    cell, src/geom/mesh.ipp:60-113): every slab addresses its planes inside them"""
    case = case_tlinear(16)
    n, hl = 16, 2
    full = n + 2 * hl + 1
    sys_full = np.full((full, full, full, 8), np.nan)
    sys_full[hl:hl + n, hl:hl + n, hl:hl + n] = case["system"]
    x_full = np.full((full, full, full), 777.0)
    conf = Conf(tol=1e-8, miniter=0, maxiter=1000)
    solver = SolverConjugateCudaGroup(conf, {}, Mesh(shape=(n, n, n), periodic=case["periodic"]),
                                      [0, 0, 0])
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


@pytest.mark.parametrize("devices", device_lists()[:1] + device_lists()[2:3])
def test_group_jacobi(gpu, devices):
    """SolverJacobi twin over slabs: max|u_new - u| all-reduced through the mailboxes;
    point Jacobi has no summation-order freedom, so the count is the reference's 613"""
    case = case_tlinear(32)
    conf = Conf(tol=1e-5, miniter=0, maxiter=1000)
    solver = SolverJacobiCudaGroup(conf, {}, Mesh(shape=(32, 32, 32), periodic=case["periodic"]),
                                   devices)
    x = np.zeros((32, 32, 32))
    info = solver.Solve(case["system"], None, x)
    solver.close()
    xo, it_o, res_o, _ = oracle_solve(case, tol=1e-5, miniter=0, maxiter=1000, method="jacobi")
    assert it_o == 613
    assert info.iter == it_o
    assert rel_max_abs(x, xo) < 1e-12


def test_group_through_the_factory(gpu):
    """`cuda_devices` on the reference's factory path (ModuleLinear::Make)"""
    case = case_tlinear(16)
    var = {"hypre_symm_tol": 1e-6, "hypre_symm_maxiter": 500, "cuda_device": 0}
    if capi.device_count() >= 2:
        var["cuda_devices"] = 2          # slabs on consecutive devices
    else:
        var["cuda_slabs_per_device"] = 2  # two slabs on the one GPU
    s = ModuleLinear.GetInstance("conjugate_cuda").Make(var, "symm", Mesh(shape=(16, 16, 16)))
    assert isinstance(s, SolverConjugateCudaGroup)
    x = np.zeros((16, 16, 16))
    info = s.Solve(case["system"], None, x)
    s.close()
    xo, it_o, _, _ = oracle_solve(case, tol=1e-6, miniter=0, maxiter=500)
    assert abs(info.iter - it_o) <= ITER_TOL and rel_max_abs(x, xo) <= 1e-9


def test_group_error_paths(gpu):
    m = Mesh(shape=(4, 8, 8))
    with pytest.raises(capi.AphcgError, match="cannot cut"):
        SolverConjugateCudaGroup(Conf(), {}, m, [0] * 5)
    with pytest.raises(capi.AphcgError, match="out of range"):
        SolverConjugateCudaGroup(Conf(), {}, m, [0, 99])
    s = SolverConjugateCudaGroup(Conf(), {}, m, [0, 0])
    with pytest.raises(capi.AphcgError, match="no system"):
        s.Run()
    # a failed group call leaves the group unusable, and says so
    with pytest.raises(capi.AphcgError, match="unusable"):
        s.Run()
    s.close()
