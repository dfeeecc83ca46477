"""BASELINE.json's full-size configurations through size-independent properties
(the oracle cannot finish them in seconds): operator linearity / symmetry / null
space, agreement of the recursively updated residual with the true residual b - A x,
the first iterations of the residual history against the oracle, run-to-run
determinism, and the known exact solution of the periodic case."""

from __future__ import annotations

import numpy as np
import pytest

from aphros_b200 import Conf, Mesh, SolverConjugateCuda, systems

pytestmark = pytest.mark.gpu


def res_norm(r, shape):
    return float(np.sqrt((r ** 2).sum() / systems.cell_volume(shape)))


def test_config2_256_variable_density(gpu):
    """config 2: 256^3 synthetic variable-density Poisson, density jump 1000:1, 1 GPU"""
    from oracle import cpu
    n = 256
    shape = (n, n, n)
    per = (False, False, False)
    solver = SolverConjugateCuda(Conf(tol=0.0, miniter=0, maxiter=24), {}, Mesh(shape=shape, periodic=per))
    solver.AssembleSpheres(systems.random_spheres(64, 20240601))
    system = solver.DownloadSystem()
    b = -system[..., 7]
    # 1. first 25 iterations of the history against the oracle
    solver.UploadGuess(None)
    info = solver.Run()
    hist = solver.History(info.iter)
    _, it_o, _, hist_o = cpu.solve(system, periodic=per, tol=0.0, miniter=0, maxiter=24)
    assert info.iter == it_o == 25
    np.testing.assert_allclose(hist, hist_o, rtol=1e-8)
    # 2. the reference recurrence is UNpreconditioned: on this 1000:1 bubble field it
    #    needs far more than 20000 iterations for 1e-8 (the reference's own
    #    SolverConjugate behaves the same: it is the same algorithm).
    #    2000 iterations: the recursively updated residual is still the true one.
    res0 = res_norm(b, shape)
    solver.SetConf(Conf(tol=0.0, miniter=0, maxiter=1999))
    solver.UploadGuess(None)
    info = solver.Run()
    assert info.iter == 2000
    x = solver.DownloadSolution(np.empty(shape))
    true_r = b - solver.Apply(x)
    assert abs(res_norm(true_r, shape) - info.residual) <= 1e-9 * res0
    # 3. determinism: same residual and bitwise the same solution
    solver.UploadGuess(None)
    info2 = solver.Run()
    x2 = solver.DownloadSolution(np.empty(shape))
    assert info2.residual == info.residual and np.array_equal(x, x2)
    solver.close()
    # 4. CG to 1e-8 relative residual on this system with the opt-in Jacobi
    #    preconditioner (the north star's "Jacobi-preconditioned CG loop")
    pre = SolverConjugateCuda(Conf(tol=1e-8 * res0, miniter=0, maxiter=20000), {"jacobi_precond": True},
                              Mesh(shape=shape, periodic=per))
    pre.AssembleSpheres(systems.random_spheres(64, 20240601))
    pre.UploadGuess(None)
    infop = pre.Run()
    assert infop.residual < 1e-8 * res0 and infop.iter < 20000, infop
    xp = pre.DownloadSolution(np.empty(shape))
    true_r = b - pre.Apply(xp)
    # after 10^4 iterations the recursive residual has drifted from the true one by a
    # few 1e-9 of res0 (the usual "residual gap" of CG; the reference never recomputes
    # it either, linear.ipp:102-107): the true residual still meets the tolerance x2
    assert res_norm(true_r, shape) <= 2e-8 * res0
    print("config 2: plain CG residual after 2000 iterations %.3e (res0 %.3e); "
          "Jacobi-PCG converged to 1e-8 in %d iterations" % (info.residual, res0, infop.iter))
    pre.close()


def test_config5_384_periodic(gpu):
    """config 5: 384^3 triply periodic constant-density projection solve, with the
    reference test's exact solution as right-hand side (src/test/linear/main.cpp:47-52,80-84);
    "iterations-to-tolerance vs reference": the reference's own SolverConjugate (oracle/_ref/
    ref_cg, 32^3 blocks, OpenMP) needs 942 iterations to 1e-7 x the initial residual on this
    system -- tests/golden/config5_384_periodic.npz, written by `make_golden.py large`"""
    import os
    from oracle import cpu
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                "config5_384_periodic.npz"))
    n = 384
    shape = (n, n, n)
    system, exact = systems.periodic_constant_system(n)
    b = -system[..., 7]
    res0 = res_norm(b, shape)
    conf = Conf(tol=1e-7 * res0, miniter=0, maxiter=5000)
    solver = SolverConjugateCuda(conf, {}, Mesh(shape=shape))
    x = np.zeros(shape)
    info = solver.Solve(system, None, x)
    hist = solver.History(info.iter)
    assert info.residual < conf.tol
    # the headline quantity of config 5: iterations to tolerance within +-2 of the reference's
    assert abs(conf.tol - float(gold["tol"])) <= 1e-12 * conf.tol
    assert abs(info.iter - int(gold["iter"])) <= 2, (info.iter, int(gold["iter"]))
    assert abs(info.residual - float(gold["residual"])) <= 1e-3 * float(gold["residual"])
    # ... and the solution at the fixture's 4096 sample cells (same mean: CG never changes it)
    xs = x.reshape(-1)[gold["sample_index"]]
    assert np.abs(xs - gold["sample_x"]).max() <= 1e-10 * np.abs(exact).max()
    print("config 5: %d iterations (reference %d), residual %.6e (reference %.6e), sample "
          "max-abs diff %.2e" % (info.iter, int(gold["iter"]), info.residual,
                                 float(gold["residual"]), np.abs(xs - gold["sample_x"]).max()))
    # constant diagonal: the Jacobi-preconditioned recurrence is the same iteration
    pre = SolverConjugateCuda(conf, {"jacobi_precond": True}, Mesh(shape=shape))
    xp = np.zeros(shape)
    infop = pre.Solve(system, None, xp)
    pre.close()
    assert infop.iter == info.iter
    # solves A x = A exact: same solution up to a constant
    d = (x - x.mean()) - (exact - exact.mean())
    assert np.abs(d).max() <= 1e-5 * np.abs(exact).max()
    # the first iterations against the oracle (each oracle iteration takes ~1 s here)
    _, it_o, _, hist_o = cpu.solve(system, tol=0.0, miniter=0, maxiter=7)
    np.testing.assert_allclose(hist[:8], hist_o[:8], rtol=1e-9)
    solver.close()


def test_config3_512_operator_and_recursion(gpu):
    """config 3/4 per-GPU size: 512^3 variable-density system assembled on the device"""
    n = 512
    shape = (n, n, n)
    per = (False, False, False)
    solver = SolverConjugateCuda(Conf(tol=0.0, miniter=0, maxiter=100), {}, Mesh(shape=shape, periodic=per))
    solver.AssembleSpheres(systems.random_spheres(512, 20240602))
    rng = np.random.default_rng(0)
    v = rng.standard_normal(shape)
    w = rng.standard_normal(shape)
    av, aw = solver.Apply(v), solver.Apply(w)
    # symmetry and linearity of the operator, null space of the Neumann problem
    s1, s2 = float((av * w).sum()), float((v * aw).sum())
    assert abs(s1 - s2) <= 1e-11 * float(np.abs(av * w).sum())
    comb = solver.Apply(2.5 * v - w)
    scale = np.abs(av).max()
    assert np.abs(comb - (2.5 * av - aw)).max() <= 1e-13 * scale
    del comb, aw, w
    ones = solver.Apply(np.ones(shape))
    assert np.abs(ones).max() <= 1e-15 * scale
    del ones, av, v
    # 101 iterations (the benchmark step): recursive residual == true residual
    solver.UploadGuess(None)
    info = solver.Run()
    assert info.iter == 101
    x = solver.DownloadSolution(np.empty(shape))
    ax = solver.Apply(x)
    system_rhs = -solver.DownloadSystem()[..., 7]
    true_r = system_rhs - ax
    res0 = res_norm(system_rhs, shape)
    assert abs(res_norm(true_r, shape) - info.residual) <= 1e-9 * res0
    solver.close()
