"""Command-line twin of the reference's linear-solver test `t.linear`
(src/test/linear/main.cpp:148-234) for the CUDA modules.

    python -m aphros_b200.tlinear --solver conjugate_cuda --tol 1e-5 --maxiter 1000 \
        --mesh 32 --verbose

Same arguments and the same report lines (`max_diff_exact`, `residual`, `iter`,
`time`; main.cpp:136-142).  The built-in system is the reference test's own
(periodic, resistivity 10 inside r<0.2; main.cpp:44-92).  Differences forced by
this image: HDF5 is absent, so `--system_in` / `--system_out` read and write the
system as raw little-endian float64 of shape (nz, ny, nx, 8) -- the shape of the
reference's HDF5 field 'data' (main.cpp:182-189) -- with the mesh size passed in
`--mesh` (or `--mesh_xyz NX NY NZ`); `--block` is accepted and ignored (the module
has no block structure).

`--replay PREFIX` solves a system captured from a live aphros run by the adapter
(`set string linsolver_symm_cuda_dump PREFIX`, aphros_b200/plugin/linear_conjugate_cuda.cpp):
PREFIX.txt gives the mesh, periodicity, cell volume and the Conf of the captured call (`--tol`
/ `--maxiter` given on the command line override it), PREFIX.sys the rows, PREFIX.x0 the guess;
`--sol_out FILE` writes the solution (raw float64).
"""

from __future__ import annotations

import argparse
import sys
import time

import numpy as np

from . import ModuleLinear, Mesh, systems


def parse_extra(text, var):
    """the reference's `set <type> <key> <value>` lines (src/parse/parser.h:15-36)"""
    for line in text.replace(";", "\n").splitlines():
        tok = line.split()
        if len(tok) >= 4 and tok[0] == "set":
            kind, key, val = tok[1], tok[2], " ".join(tok[3:])
            var[key] = int(val) if kind == "int" else float(val) if kind == "double" else val
    return var


def read_capture(prefix):
    """(meta dict, system (nz,ny,nx,8), guess or None) of an adapter capture"""
    meta = {}
    with open(prefix + ".txt") as f:
        for line in f:
            tok = line.split()
            if tok:
                meta[tok[0]] = tok[1:]
    nx, ny, nz = (int(meta[k][0]) for k in ("nx", "ny", "nz"))
    out = {"shape": (nz, ny, nx), "periodic": tuple(bool(int(v)) for v in meta["periodic"]),
           "cell_volume": float(meta["cell_volume"][0]), "tol": float(meta["tol"][0]),
           "miniter": int(meta["miniter"][0]), "maxiter": int(meta["maxiter"][0]),
           "name": " ".join(meta.get("name", [""]))}
    system = np.fromfile(prefix + ".sys", dtype=np.float64).reshape(out["shape"] + (8,))
    guess = None
    if int(meta["guess"][0]):
        guess = np.fromfile(prefix + ".x0", dtype=np.float64).reshape(out["shape"])
    return out, system, guess


def main_replay(args):
    meta, system, guess = read_capture(args.replay)
    var = {"hypre_symm_tol": meta["tol"] if args.tol is None else args.tol,
           "hypre_symm_maxiter": meta["maxiter"] if args.maxiter is None else args.maxiter,
           "hypre_symm_miniter": meta["miniter"]}
    parse_extra(args.extra, var)
    factory = ModuleLinear.GetInstance(args.solver)
    if factory is None:
        raise SystemExit("Solver not found: " + args.solver)
    t0 = time.perf_counter()
    solver = factory.Make(var, "symm", Mesh(shape=meta["shape"], periodic=meta["periodic"],
                                            cell_volume=meta["cell_volume"]))
    sol = np.zeros(meta["shape"]) if guess is None else guess.copy()
    info = solver.Solve(system, None if guess is None else sol, sol)
    dt = time.perf_counter() - t0
    solver.close()
    if args.sol_out:
        sol.tofile(args.sol_out)
    sys.stderr.write("linear(%s) '%s': res=%e iter=%d\n" % (args.solver, meta["name"], info.residual,
                                                             info.iter))
    print("residual=%.17g" % info.residual)
    print("iter=%d" % info.iter)
    print("time=%f" % dt)
    return 0


def main(argv=None):
    names = sorted(ModuleLinear.GetInstances())
    ap = argparse.ArgumentParser(description="Test for linear solvers.")
    ap.add_argument("--verbose", action="store_true", help="Print solver info.")
    ap.add_argument("--solver", default="conjugate_cuda", choices=names, help="Linear solver to use")
    ap.add_argument("--tol", type=float, default=None, help="Convergence tolerance (default 1e-3)")
    ap.add_argument("--maxiter", type=int, default=None, help="Maximum iterations (default 100)")
    ap.add_argument("--mesh", type=int, default=32, help="Mesh size in all directions")
    ap.add_argument("--mesh_xyz", type=int, nargs=3, default=None, metavar=("NX", "NY", "NZ"))
    ap.add_argument("--block", type=int, default=16, help="Block size (ignored)")
    ap.add_argument("--dump", action="store_true",
                    help="Dump solution, exact solution, and difference (sol/exact/diff_0000.raw)")
    ap.add_argument("--system_in", default="", help="raw float64 (nz,ny,nx,8) system to solve")
    ap.add_argument("--system_out", default="", help="write the system as raw float64 (nz,ny,nx,8)")
    ap.add_argument("--extra", default="", help="Extra configuration (commands 'set ... ')")
    ap.add_argument("--replay", default="", metavar="PREFIX",
                    help="solve a system captured by the aphros adapter (PREFIX.txt/.sys/.x0)")
    ap.add_argument("--sol_out", default="", help="with --replay: write the solution (raw float64)")
    args = ap.parse_args(argv)
    if args.replay:
        return main_replay(args)
    args.tol = 1e-3 if args.tol is None else args.tol
    args.maxiter = 100 if args.maxiter is None else args.maxiter

    nx, ny, nz = args.mesh_xyz if args.mesh_xyz else (args.mesh,) * 3
    shape = (nz, ny, nx)
    system, exact = systems.tlinear_system(None, shape=shape)
    if args.system_in:
        system = np.fromfile(args.system_in, dtype=np.float64).reshape(shape + (8,))
    if args.system_out:
        system.tofile(args.system_out)

    # the configuration t.linear builds (main.cpp:208-231)
    var = {"hypre_symm_tol": args.tol, "hypre_symm_maxiter": args.maxiter,
           "linsolver_symm_maxnorm": 0, "hypre_periodic_x": 1, "hypre_periodic_y": 1,
           "hypre_periodic_z": 1}
    parse_extra(args.extra, var)
    periodic = tuple(bool(var["hypre_periodic_" + d]) for d in "xyz")
    factory = ModuleLinear.GetInstance(args.solver)
    if factory is None:
        raise SystemExit("Solver not found: " + args.solver)
    t0 = time.perf_counter()
    solver = factory.Make(var, "symm", Mesh(shape=shape, periodic=periodic))
    sol = np.zeros(shape)
    info = solver.Solve(system, sol, sol)
    dt = time.perf_counter() - t0

    diff = sol - exact
    diff -= diff.mean()
    if args.dump:
        for name, f in (("sol", sol), ("exact", exact), ("diff", diff)):
            f.tofile("%s_0000.raw" % name)
    if args.verbose:
        # the report line of linear.ipp:119-124 (m.flags.linreport)
        sys.stderr.write("linear(%s) '': res=%e iter=%d\n" % (args.solver, info.residual, info.iter))
        print("\nmax_diff_exact=%g" % np.abs(diff).max())
        print("residual=%g" % info.residual)
        print("iter=%d" % info.iter)
        print("time=%f" % dt)
    return 0


if __name__ == "__main__":
    sys.exit(main())
