"""TEST INFRASTRUCTURE: ctypes binding of oracle/libcg_oracle.so (cg_oracle.h)
and a runner for oracle/_ref/ref_cg (the reference's own SolverConjugate built
from /root/reference/src by oracle/ref/Makefile).  Never imported by the
product package `aphros_b200`.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcg_oracle.so")
REF_DIR = os.path.join(_HERE, "_ref")
REF_CG = os.path.join(REF_DIR, "ref_cg")
REF_CG2 = REF_CG + "2"  # the same driver on MeshCartesian<double,2> (oracle/ref: make dim2)


class _Desc(ctypes.Structure):
    _fields_ = [
        ("nx", ctypes.c_long), ("ny", ctypes.c_long), ("nz", ctypes.c_long),
        ("periodic", ctypes.c_int * 3),
        ("bsx", ctypes.c_long), ("bsy", ctypes.c_long), ("bsz", ctypes.c_long),
        ("cell_volume", ctypes.c_double),
        ("tol", ctypes.c_double),
        ("miniter", ctypes.c_int), ("maxiter", ctypes.c_int), ("maxnorm", ctypes.c_int),
    ]


_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "libcg_oracle.so"], check=True,
                   stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        for name in ("cg_oracle_conjugate", "cg_oracle_jacobi", "cg_oracle_pconjugate"):
            f = getattr(_lib, name)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.POINTER(_Desc), dp, dp, dp, dp,
                          ctypes.POINTER(ctypes.c_int), dp]
        _lib.cg_oracle_apply.restype = ctypes.c_int
        _lib.cg_oracle_apply.argtypes = [ctypes.POINTER(_Desc), dp, dp, dp]
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def _desc(shape, periodic, block, cell_volume, tol, miniter, maxiter, maxnorm):
    nz, ny, nx = shape
    d = _Desc()
    d.nx, d.ny, d.nz = nx, ny, nz
    d.periodic[:] = [int(bool(p)) for p in periodic]
    b = block if block is not None else (0, 0, 0)
    if isinstance(b, int):
        b = (b, b, b)
    d.bsx, d.bsy, d.bsz = b
    d.cell_volume = cell_volume
    d.tol, d.miniter, d.maxiter, d.maxnorm = tol, miniter, maxiter, int(maxnorm)
    return d


def reference_cell_volume(shape, block=None):
    """m.GetCellSize().prod() as the reference's ROOT block computes it: the block
    spans [0, bs*h] with h = extent/max(n) (src/kernel/kernelmesh.h:29-44) and
    cell_size = span/bs per direction (src/geom/mesh.ipp:84-86), which is h only
    up to rounding; prod() multiplies x, y, z in order."""
    nz, ny, nx = shape
    b = block if block is not None else (nx, ny, nz)
    if isinstance(b, int):
        b = (b, b, b)
    b = [bb if bb > 0 else n for bb, n in zip(b, (nx, ny, nz))]
    h = 1.0 / max(shape)
    cs = [(float(bb) * h - 0.0 * h) / float(bb) for bb in b]
    return cs[0] * cs[1] * cs[2]


def solve(system, x0=None, *, periodic=(True, True, True), cell_volume=None, tol=0.0,
          miniter=0, maxiter=100, maxnorm=False, block=None, method="conjugate"):
    """Run the C restatement.  system: (nz,ny,nx,8).  Returns (x, iter, residual, history)."""
    system = np.ascontiguousarray(system, dtype=np.float64)
    shape = system.shape[:3]
    if cell_volume is None:
        cell_volume = reference_cell_volume(shape, block)
    x0c = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
    x = np.empty(shape, dtype=np.float64)
    hist = np.zeros(max(maxiter, miniter) + 2, dtype=np.float64)
    res = ctypes.c_double()
    it = ctypes.c_int()
    d = _desc(shape, periodic, block, cell_volume, tol, miniter, maxiter, maxnorm)
    fn = {"conjugate": lib().cg_oracle_conjugate, "jacobi": lib().cg_oracle_jacobi,
          "pconjugate": lib().cg_oracle_pconjugate}[method]
    rc = fn(ctypes.byref(d), _dp(system), _dp(x0c), _dp(x), ctypes.byref(res),
            ctypes.byref(it), _dp(hist))
    if rc != 0:
        raise MemoryError("cg_oracle: allocation failed")
    return x, it.value, res.value, hist[:it.value].copy()


def apply(system, v, *, periodic=(True, True, True)):
    system = np.ascontiguousarray(system, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.empty_like(v)
    d = _desc(system.shape[:3], periodic, None, 1.0, 0.0, 0, 0, False)
    lib().cg_oracle_apply(ctypes.byref(d), _dp(system), _dp(v), _dp(out))
    return out


REF_ASSEMBLE = os.path.join(REF_DIR, "ref_assemble")


def have_reference_assembler():
    return os.path.exists(REF_ASSEMBLE)


def assemble_reference(rho, vx, vy, vz, source=None, *, dt=1e-3, periodic=(False, False, False),
                       block=None, plugin=None, solver="conjugate_cuda", tol=1e-8, maxiter=1000,
                       threads=1, env=None):
    """Rows of the projection pressure system assembled by the reference's own functions
    (oracle/_ref/ref_assemble: InterpolateHarmonic, GradientImplicit, AppendExpr in the
    sequence of Proj::GetFlux/GetFluxSum, src/solver/proj.ipp:343-383) from a cell density
    (nz,ny,nx) and face volume fluxes vx (nz,ny,nx+1), vy (nz,ny+1,nx), vz (nz+1,ny,nx).
    Mesh extent 1 (h = 1/max(n)); non-periodic domain faces are walls.  Returns (nz,ny,nx,8).

    With `plugin` (the adapter .so) the module `solver` is also asked for the
    linear::ProjectionSolver capability and solves the same projection from the same fields,
    zero guess: returns (rows, x, iter, residual)."""
    rho = np.ascontiguousarray(rho, dtype=np.float64)
    nz, ny, nx = rho.shape
    b = block if block is not None else (nx, ny, nz)
    if isinstance(b, int):
        b = (b, b, b)
    with tempfile.TemporaryDirectory() as tmp:
        files = {}
        for name, a, shp in (("rho", rho, (nz, ny, nx)), ("vx", vx, (nz, ny, nx + 1)),
                             ("vy", vy, (nz, ny + 1, nx)), ("vz", vz, (nz + 1, ny, nx)),
                             ("src", source, (nz, ny, nx))):
            if a is None:
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != shp:
                raise ValueError("%s: expected shape %s, got %s" % (name, shp, a.shape))
            files[name] = os.path.join(tmp, name + ".f64")
            a.tofile(files[name])
        out = os.path.join(tmp, "rows.f64")
        cmd = [REF_ASSEMBLE, "--nx", str(nx), "--ny", str(ny), "--nz", str(nz),
               "--bsx", str(b[0]), "--bsy", str(b[1]), "--bsz", str(b[2]), "--dt", repr(float(dt)),
               "--px", str(int(periodic[0])), "--py", str(int(periodic[1])),
               "--pz", str(int(periodic[2])), "--out", out]
        for name, path in files.items():
            cmd += ["--" + name, path]
        xout = os.path.join(tmp, "x.f64")
        if plugin:
            cmd += ["--plugin", plugin, "--solver", solver, "--xout", xout, "--tol", repr(float(tol)),
                    "--maxiter", str(maxiter)]
        e = dict(os.environ, OMP_NUM_THREADS=str(threads))
        e.update(env or {})
        p = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True, env=e)
        if p.returncode != 0:
            raise RuntimeError("ref_assemble failed:\n" + p.stdout + p.stderr)
        rows = np.fromfile(out, dtype=np.float64).reshape(nz, ny, nx, 8)
        if not plugin:
            return rows
        x = np.fromfile(xout, dtype=np.float64).reshape(nz, ny, nx)
        with open(xout + ".info") as f:
            it, res = f.read().split()
        return rows, x, int(it), float(res)


def have_reference():
    return os.path.exists(REF_CG)


def have_reference_dim2():
    return os.path.exists(REF_CG2)


def solve_reference(system, x0=None, *, periodic=(True, True, True), tol=0.0, miniter=0,
                    maxiter=100, maxnorm=False, block=None, solver="conjugate",
                    plugin=None, threads=1, repeat=1, extra="", env=None, workdir=None, dim=3,
                    binary=None):
    """Run the reference's own solver (oracle/_ref/ref_cg; dim=2: ref_cg2, a 2-D mesh,
    system shape (1, ny, nx, 8) whose z coefficients are ignored).

    `binary`: another build of the same driver (oracle/_ref/ref_cg_native: the reference's
    default -march=native flags, used for timing only).

    Returns (x, iter, residual, seconds).  Mesh extent is 1 (h = 1/max(n)).
    """
    system = np.ascontiguousarray(system, dtype=np.float64)
    nz, ny, nx = system.shape[:3]
    if dim == 2 and nz != 1:
        raise ValueError("a 2-D system has shape (1, ny, nx, 8)")
    b = block if block is not None else (nx, ny, nz)
    if isinstance(b, int):
        b = (b, b, b)
    with tempfile.TemporaryDirectory(dir=workdir) as tmp:
        fsys = os.path.join(tmp, "sys.f64")
        system.tofile(fsys)
        cmd = [binary or (REF_CG2 if dim == 2 else REF_CG), "--nx", str(nx), "--ny", str(ny), "--nz", str(nz),
               "--bsx", str(b[0]), "--bsy", str(b[1]), "--bsz", str(b[2]),
               "--sys", fsys, "--out", os.path.join(tmp, "out"),
               "--tol", repr(float(tol)), "--maxiter", str(maxiter), "--miniter", str(miniter),
               "--maxnorm", str(int(maxnorm)),
               "--px", str(int(periodic[0])), "--py", str(int(periodic[1])),
               "--pz", str(int(periodic[2])), "--solver", solver, "--repeat", str(repeat)]
        if x0 is not None:
            fx0 = os.path.join(tmp, "x0.f64")
            np.ascontiguousarray(x0, dtype=np.float64).tofile(fx0)
            cmd += ["--x0", fx0]
        if plugin:
            cmd += ["--plugin", plugin]
        if extra:
            cmd += ["--extra", extra]
        e = dict(os.environ)
        e["OMP_NUM_THREADS"] = str(threads)
        if env:
            e.update(env)
        p = subprocess.run(cmd, cwd=tmp, env=e, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("ref_cg failed:\n" + p.stdout + p.stderr)
        x = np.fromfile(os.path.join(tmp, "out.x"), dtype=np.float64).reshape(nz, ny, nx)
        with open(os.path.join(tmp, "out.info")) as f:
            it, res, sec = f.read().split()
    return x, int(it), float(res), float(sec)
