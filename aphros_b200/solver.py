"""Host-side mirror of the reference's linear-solver interface over libaphcg.so.

Names, argument meaning and error behaviour follow the reference so that the
parity tests read like its own (src/test/linear/main.cpp):

  linear::Solver<M>            -> Solver          (Conf, Info, Solve, SetConf, GetConf;
                                                   src/linear/linear.h:15-57)
  linear::ModuleLinear<M>      -> ModuleLinear    (GetInstance(name).Make(var, prefix, m),
                                                   GetConf(var, prefix); linear.h:59-76)
  module "conjugate_cuda"      -> SolverConjugateCuda  (sibling of "conjugate",
                                                   src/linear/linear.ipp:239-253)
  module "jacobi_cuda"         -> SolverJacobiCuda     (sibling of "jacobi")

The real drop-in for aphros itself is the C++ adapter
aphros_b200/plugin/linear_conjugate_cuda.cpp; this module is the same thing for
Python callers (tests, bench).  All arithmetic happens in the CUDA library; a
missing library or GPU raises, it never falls back to the CPU.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import capi


@dataclass
class Conf:
    """linear::Solver<M>::Conf (src/linear/linear.h:21-25)."""
    tol: float = 0.0
    miniter: int = 1
    maxiter: int = 100


@dataclass
class Info:
    """linear::Solver<M>::Info (src/linear/linear.h:27-30) + device timings."""
    residual: float = 0.0
    iter: int = 0
    loop_ms: float = 0.0
    total_ms: float = 0.0
    residual0: float = 0.0  # norm of the initial residual (CG runs)


@dataclass
class Mesh:
    """What the solver reads from MeshCartesian (src/geom/mesh.h): global size,
    periodicity flags (m.flags.is_periodic) and cell volume, plus the z-slab of
    this rank when the domain is split over several GPUs."""
    shape: tuple  # (nz, ny, nx) global inner cells
    periodic: tuple = (True, True, True)  # (x, y, z)
    cell_volume: float | None = None
    rank: int = 0
    nranks: int = 1
    z0: int = 0
    nz_local: int | None = None
    device: int = 0

    def __post_init__(self):
        if self.cell_volume is None:
            from .systems import cell_volume
            self.cell_volume = cell_volume(self.shape)
        if self.nz_local is None:
            self.nz_local = self.shape[0]

    @property
    def local_shape(self):
        return (self.nz_local, self.shape[1], self.shape[2])


class Solver:
    """Abstract linear::Solver<M> (src/linear/linear.h:15-57)."""

    def __init__(self, conf: Conf):
        self.conf = conf

    def Solve(self, fc_system, fc_init, fc_sol):
        raise NotImplementedError

    def SetConf(self, c: Conf):
        self.conf = c

    def GetConf(self) -> Conf:
        return self.conf


class _CudaSolverBase(Solver):
    _method = "conjugate"

    def __init__(self, conf: Conf, extra: dict | None, m: Mesh, flags: int = 0):
        super().__init__(conf)
        extra = extra or {}
        self.mesh = m
        nz, ny, nx = m.shape
        d = capi.Desc()
        d.nx, d.ny, d.nz = nx, ny, nz
        d.periodic[:] = [int(bool(p)) for p in m.periodic]
        d.cell_volume = m.cell_volume
        d.device = m.device
        d.rank, d.nranks = m.rank, m.nranks
        d.z0, d.nz_local = m.z0, m.nz_local
        d.flags = flags | (capi.APHCG_MAXNORM if extra.get("residual_max") else 0)
        if extra.get("jacobi_precond"):
            d.flags |= capi.APHCG_JACOBI_PRECOND  # opt-in, not the reference's recurrence
        self._h = ctypes.c_void_p()
        capi.check(capi.lib().aphcg_create(ctypes.byref(self._h), ctypes.byref(d)))

    # -- lifetime ---------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            capi.lib().aphcg_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ------------------------------------------------------------------
    def _conf(self):
        return capi.Conf(float(self.conf.tol), int(self.conf.miniter), int(self.conf.maxiter))

    def _check_field(self, a, rows):
        want = self.mesh.local_shape + ((8,) if rows else ())
        if a.dtype != np.float64 or tuple(a.shape) != want:
            raise ValueError("expected float64 array of shape %s, got %s %s"
                             % (want, a.dtype, a.shape))

    def _run(self):
        info = capi.Info()
        c = self._conf()
        fn = capi.lib().aphcg_run if self._method == "conjugate" else capi.lib().aphcg_run_jacobi
        capi.check(fn(self._h, ctypes.byref(c), ctypes.byref(info)))
        return Info(info.residual, info.iter, info.loop_ms, info.total_ms, info.residual0)

    # -- linear::Solver<M>::Solve ---------------------------------------------------
    def Solve(self, fc_system, fc_init, fc_sol):
        """fc_system: (nz_local, ny, nx, 8) rows [c,x-,x+,y-,y+,z-,z+,const];
        fc_init: initial guess, may be fc_sol itself, None = zero guess
        (src/linear/linear.h:34-44); fc_sol: output array.  Arrays may be strided
        views (e.g. the inner part of a field with halos).  Returns Info."""
        L = capi.lib()
        ls = self.mesh.local_shape
        self._check_field(fc_system, True)
        self._check_field(fc_sol, False)
        lay_s = capi.layout_of(fc_system, ls, 8)
        lay_x = capi.layout_of(fc_sol, ls)
        if self._method != "conjugate":
            capi.check(L.aphcg_upload_system(self._h, capi.ptr(fc_system), ctypes.byref(lay_s)))
            self._upload_guess(fc_init)
            info = self._run()
            capi.check(L.aphcg_download_solution(self._h, capi.ptr(fc_sol), ctypes.byref(lay_x)))
            return info
        info = capi.Info()
        c = self._conf()
        if fc_init is not None:
            self._check_field(fc_init, False)
            lay_0 = capi.layout_of(fc_init, ls)
            p0, pl0 = capi.ptr(fc_init), ctypes.byref(lay_0)
        else:
            p0, pl0 = None, None
        capi.check(L.aphcg_solve(self._h, capi.ptr(fc_system), ctypes.byref(lay_s), p0, pl0,
                                 capi.ptr(fc_sol), ctypes.byref(lay_x), ctypes.byref(c),
                                 ctypes.byref(info)))
        return Info(info.residual, info.iter, info.loop_ms, info.total_ms, info.residual0)

    # -- device-resident pieces (bench, tests) ----------------------------------------
    def _upload_guess(self, fc_init):
        if fc_init is None:
            capi.check(capi.lib().aphcg_upload_guess(self._h, None, None))
        else:
            self._check_field(fc_init, False)
            lay = capi.layout_of(fc_init, self.mesh.local_shape)
            capi.check(capi.lib().aphcg_upload_guess(self._h, capi.ptr(fc_init), ctypes.byref(lay)))

    def UploadSystem(self, fc_system):
        self._check_field(fc_system, True)
        lay = capi.layout_of(fc_system, self.mesh.local_shape, 8)
        capi.check(capi.lib().aphcg_upload_system(self._h, capi.ptr(fc_system), ctypes.byref(lay)))

    def UploadGuess(self, fc_init):
        self._upload_guess(fc_init)

    def Run(self) -> Info:
        return self._run()

    def DownloadSolution(self, fc_sol):
        self._check_field(fc_sol, False)
        lay = capi.layout_of(fc_sol, self.mesh.local_shape)
        capi.check(capi.lib().aphcg_download_solution(self._h, capi.ptr(fc_sol), ctypes.byref(lay)))
        return fc_sol

    def History(self, n=None):
        n = int(n if n is not None else self.conf.maxiter + 2)
        out = np.zeros(max(n, 1), dtype=np.float64)
        got = capi.check(capi.lib().aphcg_get_history(self._h, capi.ptr(out), n))
        return out[:got]

    def Apply(self, v):
        """A*v for the resident system (stage "iter" operator, linear.ipp:65-72)."""
        self._check_field(v, False)
        out = np.empty(self.mesh.local_shape, dtype=np.float64)
        lv = capi.layout_of(v, self.mesh.local_shape)
        capi.check(capi.lib().aphcg_apply(self._h, capi.ptr(v), ctypes.byref(lv), capi.ptr(out), None))
        return out

    def TrueResidualSum(self):
        """this rank's sum r^2 with r = -(A x + e7) recomputed on the device from the
        resident solution (collective when nranks > 1; the caller adds the ranks)"""
        out = ctypes.c_double()
        capi.check(capi.lib().aphcg_true_residual(self._h, ctypes.byref(out)))
        return out.value

    def AssembleSpheres(self, spheres, rho_in=1e-3, rho_out=1.0, dt=1e-3):
        sph = np.ascontiguousarray(spheres, dtype=np.float64).reshape(-1, 4)
        capi.check(capi.lib().aphcg_assemble_spheres(self._h, capi.ptr(sph), sph.shape[0],
                                                     rho_in, rho_out, dt))

    def AssembleProjection(self, rho, vx, vy, vz, source=None, dt=1e-3, h=None):
        """device-side assembly from cell density (with z ghost planes) and face fluxes"""
        nzl, ny, nx = self.mesh.local_shape
        arrs = []
        for a, shp in ((rho, (nzl + 2, ny, nx)), (vx, (nzl, ny, nx + 1)), (vy, (nzl, ny + 1, nx)),
                       (vz, (nzl + 1, ny, nx))):
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != shp:
                raise ValueError("expected shape %s, got %s" % (shp, a.shape))
            arrs.append(a)
        src = None if source is None else np.ascontiguousarray(source, dtype=np.float64)
        h = h if h is not None else 1.0 / max(self.mesh.shape)
        capi.check(capi.lib().aphcg_assemble_projection(
            self._h, capi.ptr(arrs[0]), capi.ptr(arrs[1]), capi.ptr(arrs[2]), capi.ptr(arrs[3]),
            capi.ptr(src), dt, h))

    def DownloadSystem(self):
        out = np.empty(self.mesh.local_shape + (8,), dtype=np.float64)
        capi.check(capi.lib().aphcg_download_system(self._h, capi.ptr(out), None))
        return out

    def TimerStart(self):
        capi.check(capi.lib().aphcg_timer_start(self._h))

    def TimerStop(self) -> float:
        ms = ctypes.c_double()
        capi.check(capi.lib().aphcg_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def ProfileKernels(self, iters=20):
        a, b = ctypes.c_double(), ctypes.c_double()
        capi.check(capi.lib().aphcg_profile_kernels(self._h, iters, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def Describe(self) -> str:
        buf = ctypes.create_string_buffer(512)
        capi.check(capi.lib().aphcg_describe(self._h, buf, 512))
        return buf.value.decode()

    def LaunchCount(self):
        return int(capi.lib().aphcg_launch_count(self._h))

    def LaunchesPerIter(self):
        return int(capi.lib().aphcg_launches_per_iter(self._h))

    # -- multi-GPU wiring (see aphros_b200/distr.py) -----------------------------------
    def CommInit(self, unique_id: bytes):
        buf = ctypes.create_string_buffer(bytes(unique_id), capi.UNIQUE_ID_BYTES)
        capi.check(capi.lib().aphcg_comm_init(self._h, buf))

    def IpcExport(self) -> bytes:
        buf = ctypes.create_string_buffer(capi.IPC_BYTES)
        capi.check(capi.lib().aphcg_ipc_export(self._h, buf))
        return buf.raw

    def IpcConnect(self, blobs):
        """blobs: every rank's IpcExport() result, ordered by rank"""
        raw = b"".join(bytes(b).ljust(capi.IPC_BYTES, b"\0") for b in blobs)
        buf = ctypes.create_string_buffer(raw, len(raw))
        capi.check(capi.lib().aphcg_ipc_connect(self._h, buf, len(blobs)))


class SolverConjugateCuda(_CudaSolverBase):
    """B200 sibling of linear::SolverConjugate (src/linear/linear.h:78-100):
    unpreconditioned CG with the reference's recurrence, constants and exit rule
    (src/linear/linear.ipp:42-125)."""
    _method = "conjugate"


class SolverJacobiCuda(_CudaSolverBase):
    """B200 sibling of linear::SolverJacobi (src/linear/linear.ipp:152-237)."""
    _method = "jacobi"


class _GroupSolverBase(_CudaSolverBase):
    """The same solver over an in-process slab group (aphcg_group_*, include/aphcg.h):
    ONE process, ``devices`` = one CUDA ordinal per z-slab; arrays are rank-wide.
    What the aphros adapter uses when ``cuda_devices`` > 1."""

    def __init__(self, conf: Conf, extra: dict | None, m: Mesh, devices, flags: int = 0):
        Solver.__init__(self, conf)
        extra = extra or {}
        if m.nranks != 1:
            raise ValueError("a slab group takes the rank-wide mesh (nranks == 1)")
        self.mesh = m
        self.devices = [int(d) for d in devices]
        nz, ny, nx = m.shape
        d = capi.Desc()
        d.nx, d.ny, d.nz = nx, ny, nz
        d.periodic[:] = [int(bool(p)) for p in m.periodic]
        d.cell_volume = m.cell_volume
        d.nranks, d.nz_local = 1, nz
        d.flags = flags | (capi.APHCG_MAXNORM if extra.get("residual_max") else 0)
        if extra.get("jacobi_precond"):
            d.flags |= capi.APHCG_JACOBI_PRECOND
        dev = (ctypes.c_int32 * len(self.devices))(*self.devices)
        self._g = ctypes.c_void_p()
        capi.check(capi.lib().aphcg_group_create(ctypes.byref(self._g), ctypes.byref(d), dev,
                                                 len(self.devices)))
        self._h = None

    def close(self):
        if getattr(self, "_g", None) is not None and self._g.value:
            capi.lib().aphcg_group_destroy(self._g)
            self._g = ctypes.c_void_p()

    def _member(self, q=0):
        return ctypes.c_void_p(capi.lib().aphcg_group_member(self._g, q))

    def Slabs(self):
        out = []
        for q in range(len(self.devices)):
            z0, nzl = ctypes.c_int64(), ctypes.c_int64()
            capi.check(capi.lib().aphcg_group_slab(self._g, q, ctypes.byref(z0), ctypes.byref(nzl)))
            out.append((z0.value, nzl.value))
        return out

    def _run(self):
        info = capi.Info()
        c = self._conf()
        L = capi.lib()
        fn = L.aphcg_group_run if self._method == "conjugate" else L.aphcg_group_run_jacobi
        capi.check(fn(self._g, ctypes.byref(c), ctypes.byref(info)))
        return Info(info.residual, info.iter, info.loop_ms, info.total_ms, info.residual0)

    def Solve(self, fc_system, fc_init, fc_sol):
        L = capi.lib()
        ls = self.mesh.local_shape
        self._check_field(fc_system, True)
        self._check_field(fc_sol, False)
        lay_s = capi.layout_of(fc_system, ls, 8)
        lay_x = capi.layout_of(fc_sol, ls)
        if self._method != "conjugate":
            self.UploadSystem(fc_system)
            self._upload_guess(fc_init)
            info = self._run()
            self.DownloadSolution(fc_sol)
            return info
        info = capi.Info()
        c = self._conf()
        if fc_init is not None:
            self._check_field(fc_init, False)
            lay_0 = capi.layout_of(fc_init, ls)
            p0, pl0 = capi.ptr(fc_init), ctypes.byref(lay_0)
        else:
            p0, pl0 = None, None
        capi.check(L.aphcg_group_solve(self._g, capi.ptr(fc_system), ctypes.byref(lay_s), p0, pl0,
                                       capi.ptr(fc_sol), ctypes.byref(lay_x), ctypes.byref(c),
                                       ctypes.byref(info)))
        return Info(info.residual, info.iter, info.loop_ms, info.total_ms, info.residual0)

    def _upload_guess(self, fc_init):
        if fc_init is None:
            capi.check(capi.lib().aphcg_group_upload_guess(self._g, None, None))
        else:
            self._check_field(fc_init, False)
            lay = capi.layout_of(fc_init, self.mesh.local_shape)
            capi.check(capi.lib().aphcg_group_upload_guess(self._g, capi.ptr(fc_init),
                                                           ctypes.byref(lay)))

    def UploadSystem(self, fc_system):
        self._check_field(fc_system, True)
        lay = capi.layout_of(fc_system, self.mesh.local_shape, 8)
        capi.check(capi.lib().aphcg_group_upload_system(self._g, capi.ptr(fc_system),
                                                        ctypes.byref(lay)))

    def DownloadSolution(self, fc_sol):
        self._check_field(fc_sol, False)
        lay = capi.layout_of(fc_sol, self.mesh.local_shape)
        capi.check(capi.lib().aphcg_group_download_solution(self._g, capi.ptr(fc_sol),
                                                            ctypes.byref(lay)))
        return fc_sol

    def AssembleSpheres(self, spheres, rho_in=1e-3, rho_out=1.0, dt=1e-3):
        sph = np.ascontiguousarray(spheres, dtype=np.float64).reshape(-1, 4)
        capi.check(capi.lib().aphcg_group_assemble_spheres(self._g, capi.ptr(sph), sph.shape[0],
                                                           rho_in, rho_out, dt))

    def TrueResidualSum(self):
        """sum r^2 over all slabs, r = -(A x + e7) recomputed from the resident solution"""
        out = ctypes.c_double()
        capi.check(capi.lib().aphcg_group_true_residual(self._g, ctypes.byref(out)))
        return out.value

    def History(self, n=None):
        n = int(n if n is not None else self.conf.maxiter + 2)
        out = np.zeros(max(n, 1), dtype=np.float64)
        got = capi.check(capi.lib().aphcg_get_history(self._member(0), capi.ptr(out), n))
        return out[:got]

    def Describe(self) -> str:
        buf = ctypes.create_string_buffer(512)
        capi.check(capi.lib().aphcg_describe(self._member(0), buf, 512))
        return buf.value.decode() + " slabs=%d" % len(self.devices)

    def LaunchCount(self):
        return sum(int(capi.lib().aphcg_launch_count(self._member(q)))
                   for q in range(len(self.devices)))

    def LaunchesPerIter(self):
        return int(capi.lib().aphcg_launches_per_iter(self._member(0)))

    def TimerStart(self):
        for q in range(len(self.devices)):
            capi.check(capi.lib().aphcg_timer_start(self._member(q)))

    def TimerStop(self) -> float:
        worst = 0.0
        for q in range(len(self.devices)):
            ms = ctypes.c_double()
            capi.check(capi.lib().aphcg_timer_stop(self._member(q), ctypes.byref(ms)))
            worst = max(worst, ms.value)
        return worst

    def _unsupported(self, *a, **k):
        raise NotImplementedError("not available on a slab group; use the slab handles")

    Apply = AssembleProjection = DownloadSystem = ProfileKernels = _unsupported
    CommInit = IpcExport = IpcConnect = _unsupported


class SolverConjugateCudaGroup(_GroupSolverBase):
    """SolverConjugateCuda over several GPUs of one process."""
    _method = "conjugate"


class SolverJacobiCudaGroup(_GroupSolverBase):
    _method = "jacobi"


class ModuleLinear:
    """linear::ModuleLinear<M> registry (src/linear/linear.h:59-76,
    src/util/module.h:21-85)."""
    _instances: dict = {}

    def __init__(self, name):
        self.name = name

    @classmethod
    def Register(cls, mod):
        if mod.name in cls._instances:
            raise RuntimeError("RegisterModule: module '%s' already registered" % mod.name)
        cls._instances[mod.name] = mod
        return True

    @classmethod
    def GetInstance(cls, name):
        return cls._instances.get(name)

    @classmethod
    def GetInstances(cls):
        return dict(cls._instances)

    @staticmethod
    def GetConf(var: dict, prefix: str) -> Conf:
        # keys are hypre_<prefix>_* even for non-hypre modules (linear.h:66-75);
        # tol and maxiter are mandatory (Vars::operator[] throws), miniter defaults to 0
        p = "hypre_" + prefix + "_"
        return Conf(tol=float(var[p + "tol"]), maxiter=int(var[p + "maxiter"]),
                    miniter=int(var.get(p + "miniter", 0)))

    def Make(self, var: dict, prefix: str, m: Mesh) -> Solver:
        raise NotImplementedError


class ModuleLinearConjugateCuda(ModuleLinear):
    def __init__(self):
        super().__init__("conjugate_cuda")

    def Make(self, var, prefix, m):
        # extras follow the "linsolver_<prefix>_<key>" convention (linear.ipp:245-249);
        # read with a default: Vars["k"] throws on a missing key
        extra = {"residual_max": bool(int(var.get("linsolver_" + prefix + "_maxnorm", 0))),
                 "jacobi_precond": bool(int(var.get("linsolver_" + prefix + "_jacobi", 0)))}
        flags = 0
        if int(var.get("linsolver_" + prefix + "_cuda_graph", 1)) == 0:
            flags |= capi.APHCG_NO_GRAPH
        if int(var.get("linsolver_" + prefix + "_cuda_tma", 1)) == 0:
            flags |= capi.APHCG_NO_TMA
        if int(var.get("linsolver_" + prefix + "_cuda_persistent", 1)) == 0:
            flags |= capi.APHCG_NO_PERSISTENT
        if int(var.get("linsolver_" + prefix + "_cuda_stream", 1)) == 0:
            flags |= capi.APHCG_NO_STREAM
        ndev = int(var.get("cuda_devices", 1))
        per_dev = int(var.get("cuda_slabs_per_device", 1))
        if ndev * per_dev > 1:  # one process, several z-slabs: devices cuda_device..+ndev-1
            first = int(var.get("cuda_device", 0))
            devices = [first + i for i in range(ndev) for _ in range(per_dev)]
            return SolverConjugateCudaGroup(self.GetConf(var, prefix), extra, m, devices, flags)
        return SolverConjugateCuda(self.GetConf(var, prefix), extra, m, flags)


class ModuleLinearJacobiCuda(ModuleLinear):
    def __init__(self):
        super().__init__("jacobi_cuda")

    def Make(self, var, prefix, m):
        return SolverJacobiCuda(self.GetConf(var, prefix), {}, m, 0)


_kReg = [ModuleLinear.Register(ModuleLinearConjugateCuda()),
         ModuleLinear.Register(ModuleLinearJacobiCuda())]
