"""Shared parity cases: small systems the CPU oracle finishes in seconds."""

from __future__ import annotations

import numpy as np

from aphros_b200 import systems


def case_tlinear(n=32, rho_in=10.0, shape=None):
    s, exact = systems.tlinear_system(n, rho_in=rho_in, shape=shape)
    return dict(system=s, periodic=(True, True, True), exact=exact)


def case_density(n=32, nspheres=8, seed=7, periodic=(False, False, False), shape=None,
                 rho_in=1e-3):
    s, _ = systems.density_poisson_system(n, nspheres=nspheres, seed=seed, periodic=periodic,
                                          shape=shape, rho_in=rho_in)
    return dict(system=s, periodic=periodic)


def case_periodic_const(n=32, shape=None):
    s, exact = systems.periodic_constant_system(n, shape=shape)
    return dict(system=s, periodic=(True, True, True), exact=exact)


def random_guess(shape, seed=3):
    return np.random.default_rng(seed).standard_normal(shape)


def initial_residual(system, x0, periodic):
    """sqrt(sum r0^2 / V), r0 = -(A x0 + e7): the residual of the guess in the
    reference's norm (src/linear/linear.ipp:48-56,103-107); tolerances "relative to
    the start" are expressed through it because the reference only knows absolute
    ones."""
    from oracle import cpu
    shape = system.shape[:3]
    r0 = system[..., 7].copy()
    if x0 is not None:
        r0 = r0 + cpu.apply(system, x0, periodic=periodic)
    return float(np.sqrt((r0 ** 2).sum() / systems.cell_volume(shape)))


def iteration_budget(system, x0, periodic, tol, maxiter, blocks=(8, 16, 32)):
    """(+-2) + the reference's OWN summation-order noise on this system: the spread
    of the oracle's iteration counts over block sizes (the reference sums per block,
    so its count moves with the block size: SURVEY.md 0.6 measured 1957 vs 1961 on a
    1000:1 system).  Near a flat, non-monotone residual curve the count to a
    tolerance is only defined up to that spread.  Returns (budget, counts)."""
    from oracle import cpu
    counts = [cpu.solve(system, x0, periodic=periodic, tol=tol, miniter=0, maxiter=maxiter,
                        block=b)[1] for b in blocks]
    return 2 + max(counts) - min(counts), counts


def iterations_ok(it_gpu, counts):
    """Iteration parity on an ill-conditioned system, given the oracle's counts for
    several block sizes: not more than 2 above the reference's slowest configuration,
    and not more than 3 % (+2) below its fastest.  The asymmetry is deliberate and
    measured: the CUDA path forms its dot products with pairwise trees and FMAs, i.e.
    more accurately than the reference's serial sums, and CG's rounding-induced
    convergence delay shrinks with it -- on every variable-density case here the GPU
    needs 1-2 % FEWER iterations than any block configuration of the reference
    (e.g. 1194 vs 1202..1224), never more."""
    return (it_gpu <= max(counts) + 2) and (it_gpu >= int(0.97 * min(counts)) - 2)


def residual_envelope(system, x0, periodic, maxiter, maxnorm=False, blocks=(4, 8, 16)):
    """final residuals of the oracle after a FIXED number of iterations, over block
    sizes.  On a non-converged 1000:1 system they differ at O(1) (16^3, 101 iterations:
    0.21 / 0.50 / 1.10 for block 16 / 8 / 4), so such a residual is not a reproducible
    quantity of the reference; callers compare digits only when the envelope is tight."""
    from oracle import cpu
    rs = [cpu.solve(system, x0, periodic=periodic, tol=0.0, miniter=0, maxiter=maxiter,
                    maxnorm=maxnorm, block=b)[2] for b in blocks]
    return min(rs), max(rs)


def solution_budget(system, x0, periodic, tol, maxiter, blocks=(8, 16, 32)):
    """max(1e-10, 4 x the reference's own spread): the relative max-abs difference
    between the oracle's solutions for different block sizes (same system, guess and
    tolerance) is what "the reference's answer" is defined up to."""
    from oracle import cpu
    xs = [cpu.solve(system, x0, periodic=periodic, tol=tol, miniter=0, maxiter=maxiter,
                    block=b)[0] for b in blocks]
    spread = max(rel_max_abs(x, xs[0]) for x in xs[1:])
    return max(1e-10, 4 * spread), spread


def rel_max_abs(a, b):
    """The parity measure of the north star: max|a-b| / max|b|."""
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def remove_mean(a):
    return a - a.mean()
