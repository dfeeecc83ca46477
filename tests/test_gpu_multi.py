"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2`): z-slabs on separate
processes, residual ghost planes written into the neighbour GPU's memory over
NVLink, scalars all-reduced with NCCL; compared with the single-domain oracle."""

from __future__ import annotations

import os
import subprocess
import sys

import pytest

from aphros_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np
import torch, torch.distributed as dist
from aphros_b200 import Conf, SolverConjugateCuda, distr, systems
from cases import initial_residual, iteration_budget, iterations_ok, rel_max_abs, solution_budget
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
from oracle import cpu
ok = True
for name in ["tlinear_periodic", "density_walls", "uneven"]:
    if name == "tlinear_periodic":
        shape, per = (32, 32, 32), (True, True, True)
        s, _ = systems.tlinear_system(32)
    elif name == "density_walls":
        shape, per = (48, 32, 64), (False, False, False)
        s, _ = systems.density_poisson_system(None, nspheres=6, seed=4, rho_in=0.05, shape=shape)
    else:
        shape, per = (world * 5 + 1, 12, 130), (True, True, True)
        s, _ = systems.tlinear_system(None, shape=shape)
    x0 = np.random.default_rng(1).standard_normal(shape) * 1e-3
    tol = 1e-10 * initial_residual(s, x0, per)
    conf = Conf(tol=tol, miniter=0, maxiter=3000)
    m = distr.local_mesh(shape, per, rank, world, device=rank)
    solver = SolverConjugateCuda(conf, {}, m)
    distr.connect(solver)
    sl = slice(m.z0, m.z0 + m.nz_local)
    x = np.zeros(m.local_shape)
    info = solver.Solve(np.ascontiguousarray(s[sl]), np.ascontiguousarray(x0[sl]), x)
    xo, it_o, res_o, _ = cpu.solve(s, x0, periodic=per, tol=tol, miniter=0, maxiter=3000)
    err = rel_max_abs(x, xo[sl]) * np.abs(xo[sl]).max() / np.abs(xo).max()
    # second solve on the same handle: zero guess, fixed iteration count
    solver.SetConf(Conf(tol=0.0, miniter=0, maxiter=30))
    x2 = np.zeros(m.local_shape)
    info2 = solver.Solve(np.ascontiguousarray(s[sl]), None, x2)
    xo2, it_o2, res_o2, _ = cpu.solve(s, None, periodic=per, tol=0.0, miniter=0, maxiter=30)
    # the residual of a NON-converged run on the variable-density system depends on the
    # summation order at O(1) (the oracle itself: 24.2 / 11.8 / 11.2 for one block /
    # 8^3 / 16^3 blocks), so it is compared only where the oracle is stable
    res_b = cpu.solve(s, None, periodic=per, tol=0.0, miniter=0, maxiter=30, block=8)[2]
    stable = abs(res_b - res_o2) <= 1e-9 * res_o2
    budget, counts = iteration_budget(s, x0, per, tol, 3000, blocks=(4, 8, 16))
    xbudget, spread = solution_budget(s, x0, per, tol, 3000, blocks=(4, 8, 16))
    good = (iterations_ok(info.iter, counts + [it_o]) and err <= xbudget and info2.iter == it_o2
            and (not stable or abs(info2.residual - res_o2) <= 1e-7 * res_o2))
    print("rank %%d %%s: iter %%d/%%d err %%.2e | iter %%d/%%d res %%.6e/%%.6e %%s [%%s]" %% (
        rank, name, info.iter, it_o, err, info2.iter, it_o2, info2.residual, res_o2,
        "OK" if good else "FAIL", solver.Describe().split("allreduce=")[-1]), flush=True)
    ok = ok and good
    solver.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


@pytest.mark.parametrize("world,allreduce", [(2, "mail"), (2, "mail-finish"), (2, "nccl"),
                                             (4, "mail"), (8, "mail"), (8, "mail-finish")])
def test_slabs_match_single_domain_oracle(gpu, tmp_path, world, allreduce):
    """allreduce: "mail" = scalars through peer-memory mailboxes written by the kernels and
    awaited by the consumer kernels' own CTAs (the product path: two launches per iteration);
    "mail-finish" = the same mailboxes awaited by one-warp k_finish_* kernels (what slabs that
    share a GPU use); "nccl" = ncclAllReduce on the same stream (comparison path)"""
    if capi.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1",
               MASTER_PORT=str(29600 + world + {"mail": 0, "nccl": 50, "mail-finish": 100}[allreduce]),
               WORLD_SIZE=str(world), APHCG_ALLREDUCE=allreduce.split("-")[0],
               APHCG_WAIT="finish" if allreduce == "mail-finish" else "kernel")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            pytest.fail("multi-GPU worker timed out")
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o[-3000:])
    print("".join(o for o in outs[:1]))  # rank 0's lines, for the saved logs (pytest -s / -rP)
