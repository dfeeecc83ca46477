"""ctypes binding of libaphcg.so (include/aphcg.h).  Test/bench harness plumbing:
the product is the shared library; nothing here computes.

The library is loaded from this package directory (built in-tree by
``aphros_b200.build``).  Loading never needs a GPU; every compute entry point
fails loudly (``AphcgError``) when no CUDA device is usable -- there is no CPU
fallback.
"""

from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaphcg.so")

APHCG_MAXNORM = 1 << 0
APHCG_NO_GRAPH = 1 << 1
APHCG_NO_TMA = 1 << 2
APHCG_NO_SYM = 1 << 3
APHCG_NCCL_REDUCE = 1 << 4
APHCG_JACOBI_PRECOND = 1 << 5
APHCG_NO_PERSISTENT = 1 << 6
APHCG_NO_STREAM = 1 << 7
UNIQUE_ID_BYTES = 128
IPC_BYTES = 128


class AphcgError(RuntimeError):
    pass


class Desc(ctypes.Structure):
    _fields_ = [
        ("nx", ctypes.c_int64), ("ny", ctypes.c_int64), ("nz", ctypes.c_int64),
        ("periodic", ctypes.c_int32 * 3),
        ("cell_volume", ctypes.c_double),
        ("device", ctypes.c_int32),
        ("rank", ctypes.c_int32), ("nranks", ctypes.c_int32),
        ("z0", ctypes.c_int64), ("nz_local", ctypes.c_int64),
        ("flags", ctypes.c_uint32),
    ]


class Layout(ctypes.Structure):
    _fields_ = [("offset", ctypes.c_int64), ("stride_y", ctypes.c_int64),
                ("stride_z", ctypes.c_int64)]


class Conf(ctypes.Structure):
    _fields_ = [("tol", ctypes.c_double), ("miniter", ctypes.c_int32),
                ("maxiter", ctypes.c_int32)]


class Info(ctypes.Structure):
    _fields_ = [("residual", ctypes.c_double), ("iter", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("loop_ms", ctypes.c_double),
                ("total_ms", ctypes.c_double), ("residual0", ctypes.c_double)]


_lib = None

_VP = ctypes.c_void_p
_PL = ctypes.POINTER(Layout)

# name -> (restype, argtypes); every symbol include/aphcg.h declares
SIGNATURES = {
    "aphcg_last_error": (ctypes.c_char_p, []),
    "aphcg_version": (ctypes.c_int, []),
    "aphcg_device_count": (ctypes.c_int, []),
    "aphcg_create": (ctypes.c_int, [ctypes.POINTER(_VP), ctypes.POINTER(Desc)]),
    "aphcg_destroy": (ctypes.c_int, [_VP]),
    "aphcg_host_alloc": (ctypes.c_int, [ctypes.POINTER(_VP), ctypes.c_uint64]),
    "aphcg_host_free": (ctypes.c_int, [_VP]),
    "aphcg_solve": (ctypes.c_int, [_VP, _VP, _PL, _VP, _PL, _VP, _PL, ctypes.POINTER(Conf),
                                   ctypes.POINTER(Info)]),
    "aphcg_upload_system": (ctypes.c_int, [_VP, _VP, _PL]),
    "aphcg_upload_guess": (ctypes.c_int, [_VP, _VP, _PL]),
    "aphcg_set_system_device": (ctypes.c_int, [_VP, _VP, _PL]),
    "aphcg_set_guess_device": (ctypes.c_int, [_VP, _VP, _PL]),
    "aphcg_run": (ctypes.c_int, [_VP, ctypes.POINTER(Conf), ctypes.POINTER(Info)]),
    "aphcg_download_solution": (ctypes.c_int, [_VP, _VP, _PL]),
    "aphcg_get_solution_device": (ctypes.c_int, [_VP, _VP, _PL]),
    "aphcg_get_history": (ctypes.c_int, [_VP, _VP, ctypes.c_int32]),
    "aphcg_run_jacobi": (ctypes.c_int, [_VP, ctypes.POINTER(Conf), ctypes.POINTER(Info)]),
    "aphcg_apply": (ctypes.c_int, [_VP, _VP, _PL, _VP, _PL]),
    "aphcg_group_assemble_projection": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, ctypes.c_double,
                                                       ctypes.c_double]),
    "aphcg_true_residual": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_double)]),
    "aphcg_group_true_residual": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_double)]),
    "aphcg_assemble_spheres": (ctypes.c_int, [_VP, _VP, ctypes.c_int32, ctypes.c_double,
                                              ctypes.c_double, ctypes.c_double]),
    "aphcg_assemble_projection": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, ctypes.c_double,
                                                 ctypes.c_double]),
    "aphcg_download_system": (ctypes.c_int, [_VP, _VP, _PL]),
    "aphcg_comm_unique_id": (ctypes.c_int, [_VP]),
    "aphcg_comm_init": (ctypes.c_int, [_VP, _VP]),
    "aphcg_ipc_export": (ctypes.c_int, [_VP, _VP]),
    "aphcg_ipc_connect": (ctypes.c_int, [_VP, _VP, ctypes.c_int32]),
    "aphcg_group_create": (ctypes.c_int, [ctypes.POINTER(_VP), ctypes.POINTER(Desc),
                                          ctypes.POINTER(ctypes.c_int32), ctypes.c_int32]),
    "aphcg_group_destroy": (ctypes.c_int, [_VP]),
    "aphcg_group_size": (ctypes.c_int, [_VP]),
    "aphcg_group_member": (_VP, [_VP, ctypes.c_int32]),
    "aphcg_group_slab": (ctypes.c_int, [_VP, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64),
                                        ctypes.POINTER(ctypes.c_int64)]),
    "aphcg_group_solve": (ctypes.c_int, [_VP, _VP, _PL, _VP, _PL, _VP, _PL, ctypes.POINTER(Conf),
                                         ctypes.POINTER(Info)]),
    "aphcg_group_upload_system": (ctypes.c_int, [_VP, _VP, _PL]),
    "aphcg_group_upload_guess": (ctypes.c_int, [_VP, _VP, _PL]),
    "aphcg_group_run": (ctypes.c_int, [_VP, ctypes.POINTER(Conf), ctypes.POINTER(Info)]),
    "aphcg_group_run_jacobi": (ctypes.c_int, [_VP, ctypes.POINTER(Conf), ctypes.POINTER(Info)]),
    "aphcg_group_download_solution": (ctypes.c_int, [_VP, _VP, _PL]),
    "aphcg_group_assemble_spheres": (ctypes.c_int, [_VP, _VP, ctypes.c_int32, ctypes.c_double,
                                                    ctypes.c_double, ctypes.c_double]),
    "aphcg_timer_start": (ctypes.c_int, [_VP]),
    "aphcg_timer_stop": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_double)]),
    "aphcg_profile_kernels": (ctypes.c_int, [_VP, ctypes.c_int32, ctypes.POINTER(ctypes.c_double),
                                             ctypes.POINTER(ctypes.c_double)]),
    "aphcg_describe": (ctypes.c_int, [_VP, ctypes.c_char_p, ctypes.c_int32]),
    "aphcg_stream": (_VP, [_VP]),
    "aphcg_launch_count": (ctypes.c_int64, [_VP]),
    "aphcg_launches_per_iter": (ctypes.c_int, [_VP]),
}


def lib():
    """Loads libaphcg.so (building it first if the sources are newer)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(rc):
    if rc < 0:
        raise AphcgError("aphcg error %d: %s" % (rc, lib().aphcg_last_error().decode()))
    return rc


def device_count():
    return lib().aphcg_device_count()


def ptr(a):
    """Address of a numpy array / int / None as c_void_p."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return ctypes.c_void_p(int(a))
    return ctypes.c_void_p(a.ctypes.data)


def layout_of(arr, shape, row_doubles=1):
    """aphcg_layout of a numpy array whose leading axes are (nz, ny, nx)."""
    nz, ny, nx = shape
    it = arr.itemsize * row_doubles
    s = arr.strides
    if s[2] != it:
        raise ValueError("x must be contiguous")
    if s[1] % it or s[0] % it:
        raise ValueError("strides must be whole cells")
    return Layout(0, s[1] // it, s[0] // it)


class PinnedArray:
    """float64 numpy array over cudaHostAlloc'ed memory."""

    def __init__(self, shape):
        self.nbytes = int(np.prod(shape)) * 8
        p = ctypes.c_void_p()
        check(lib().aphcg_host_alloc(ctypes.byref(p), max(self.nbytes, 8)))
        self._ptr = p
        buf = (ctypes.c_double * (self.nbytes // 8)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=np.float64).reshape(shape)

    def free(self):
        if self._ptr is not None:
            self.array = None
            lib().aphcg_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
