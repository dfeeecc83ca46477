"""A few small solves through every loop implementation, for compute-sanitizer
(`compute-sanitizer --tool memcheck python scripts/sanitize_small.py`): the stream kernel
(periodic, walls with ragged sizes, periodic z), the LDG-fed and 7-stream kernels, the plain
kernel, the persistent cooperative loop, device-side assembly.  Single GPU, no slab groups
(their spinning kernels must not be serialised)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from aphros_b200 import Conf, Mesh, SolverConjugateCuda, capi, systems  # noqa: E402
from cases import case_density, case_tlinear  # noqa: E402

os.environ["APHCG_PERSISTENT"] = "0"
cases = [case_tlinear(24), case_density(None, shape=(10, 21, 70), rho_in=0.1, nspheres=3),
         case_density(None, shape=(9, 16, 64), rho_in=0.1, nspheres=2, periodic=(False, False, True))]
for case in cases:
    shape = case["system"].shape[:3]
    ref = None
    for flags in (0, capi.APHCG_NO_STREAM, capi.APHCG_NO_SYM, capi.APHCG_NO_TMA):
        s = SolverConjugateCuda(Conf(tol=0.0, miniter=0, maxiter=11), {},
                                Mesh(shape=shape, periodic=case["periodic"]), flags)
        x = np.zeros(shape)
        info = s.Solve(case["system"], None, x)
        tag = s.Describe().split(" ")[0]
        s.close()
        ref = x if ref is None else ref
        print(shape, tag, info.iter, "%.6e" % info.residual, "%.2e" % (np.abs(x - ref).max() / np.abs(ref).max()))
os.environ["APHCG_PERSISTENT"] = "1"
case = case_tlinear(16)
s = SolverConjugateCuda(Conf(tol=0.0, miniter=0, maxiter=11), {}, Mesh(shape=(16, 16, 16), periodic=case["periodic"]))
x = np.zeros((16, 16, 16))
print(s.Describe().split(" ")[0], s.Solve(case["system"], None, x).iter)
s.close()
os.environ["APHCG_PERSISTENT"] = "0"
shape = (6, 8, 20)
rho, vx, vy, vz = systems.projection_inputs(shape, systems.random_spheres(2, 1), rho_in=0.1)
s = SolverConjugateCuda(Conf(tol=0.0, miniter=0, maxiter=5), {}, Mesh(shape=shape, periodic=(False,) * 3))
s.AssembleProjection(rho, np.ascontiguousarray(vx), np.ascontiguousarray(vy), np.ascontiguousarray(vz), dt=1e-3)
s.UploadGuess(None)
print("projection", s.Run().iter, "%.3e" % s.TrueResidualSum())
s.close()
print("sanitize_small: done")
