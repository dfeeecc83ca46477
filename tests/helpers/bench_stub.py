"""Runs bench.py's own arm with the CUDA solver, torch.cuda and torch.distributed stubbed out,
so that the assembly of the JSON line (the driver's contract) is checked on a machine without
a GPU.  Usage: python bench_stub.py N  (N = 1 or 2); prints bench.py's line."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
world = int(sys.argv[1])
os.environ.update(WORLD_SIZE=str(world), RANK="0", LOCAL_RANK="0")

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
_tensor = torch.tensor
torch.tensor = lambda *a, **k: _tensor(*a, **{q: v for q, v in k.items() if q != "device"})
for name in ("init_process_group", "barrier", "all_reduce", "destroy_process_group"):
    setattr(dist, name, lambda *a, **k: None)

import aphros_b200  # noqa: E402
from aphros_b200 import capi, distr, solver as S  # noqa: E402

capi.device_count = lambda: world
distr.connect = lambda s: None


class StubSolver:
    def __init__(self, conf, extra, mesh, flags=0):
        self.mesh, self.conf, self._h, self.n = mesh, conf, None, 0

    def Run(self):
        return S.Info(81.99, self.conf.maxiter + 1, 256.0, 258.0, 2.72)

    def Solve(self, a, b, c):
        return self.Run()

    def LaunchCount(self):
        self.n += 205
        return self.n

    def Describe(self):
        return ("spmv=tma-sym4 tile=128x8 planes_per_cta=32 stages=3 l2_prefetch=2 pstream=1 "
                "ctas=4096 precond=none graph=1 allreduce=none")

    TimerStop = lambda self: 1290.0
    ProfileKernels = lambda self, n: (1.89, 0.525)
    AssembleSpheres = UploadGuess = TimerStart = SetConf = close = lambda self, *a, **k: None
    AssembleProjection = DownloadSolution = lambda self, *a, **k: None


aphros_b200.SolverConjugateCuda = StubSolver
import bench  # noqa: E402

bench.ClockSampler.start = lambda self: None
bench.ClockSampler.stop = lambda self: {"sm_mhz": 1800.0, "sm_max_mhz": 1965.0,
                                        "reasons": ["sw_power_cap"], "samples": 5}
bench.run_reference_cpu = lambda *a, **k: {"value": 3.9e8, "unit": bench.UNIT, "cores": 16,
                                           "kind": "reference", "sample": "stub"}
bench.parity_check = lambda world, rank, local_rank, d: {
    "rel_max_abs": 2e-13, "iter": 1730, "iter_reference": 1730, "ok": True}
bench.strong_measurement = lambda args, world, rank, local_rank, d, barrier: {
    "value": 9.0e10, "unit": bench.UNIT, "n1_value": 5.2e10, "efficiency_vs_n1": 9.0 / (world * 5.2)}
capi.PinnedArray = type("PA", (), {"__init__": lambda self, shape: setattr(self, "array", np.zeros(2)),
                                   "free": lambda self: None})
from aphros_b200 import systems  # noqa: E402
systems.projection_inputs = lambda *a, **k: tuple(np.zeros(2) for _ in range(4))
capi.lib = lambda: types.SimpleNamespace(aphcg_download_system=lambda *a: 0)
capi.ptr = lambda a: None
sys.argv = ["bench.py", "--gpus", str(world), "--steps", "2", "--warmup", "1"]
sys.exit(bench.main())
