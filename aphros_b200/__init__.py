"""aphros_b200 -- B200-native conjugate-gradient module for cselab/aphros'
pressure-Poisson systems (the `linear::Solver<M>` hot path only).

  csrc/      hand-written CUDA (sm_100a) kernels + the C ABI (include/aphcg.h)
  plugin/    the aphros-side adapter: ModuleLinear "conjugate_cuda" (C++)
  solver.py  the same interface for Python callers (tests, bench)
  distr.py   z-slab decomposition / multi-GPU wiring
  systems.py synthetic input systems
"""

from .solver import (Conf, Info, Mesh, ModuleLinear, Solver, SolverConjugateCuda,  # noqa: F401
                     SolverConjugateCudaGroup, SolverJacobiCuda, SolverJacobiCudaGroup)

__all__ = ["Conf", "Info", "Mesh", "ModuleLinear", "Solver", "SolverConjugateCuda",
           "SolverConjugateCudaGroup", "SolverJacobiCuda", "SolverJacobiCudaGroup"]
