"""Per-iteration latency on small meshes for a few kernel configurations (GPU)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from aphros_b200 import Conf, Mesh, SolverConjugateCuda, systems
for n in (32, 64, 96, 128, 192, 256):
    s, _ = systems.tlinear_system(n)
    solver = SolverConjugateCuda(Conf(tol=0.0, miniter=0, maxiter=399), {}, Mesh(shape=(n, n, n)))
    solver.UploadSystem(s)
    best = 1e9
    for rep in range(3):
        solver.UploadGuess(None)
        info = solver.Run()
        best = min(best, info.loop_ms / info.iter * 1e3)
    print("n=%%3d  %%7.2f us/iter  %%.3e cell-iter/s  %%s" %% (n, best, n**3 / best * 1e6, solver.Describe()), flush=True)
    solver.close()
''' % ROOT
for cfg in sys.argv[1:]:
    env = dict(os.environ)
    for kv in cfg.split(","):
        if "=" in kv:
            k, v = kv.split("=", 1)
            env[k] = v
    print("==", cfg, flush=True)
    subprocess.run([sys.executable, "-c", CODE], env=env)
