"""Builds libaphcg.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libaphcg.so")
SOURCES = ["aphcg.cu", "aphcg_group.cu", "cg_kernels.cu", "cg_spmv_tma.cu", "cg_spmv_tma2.cu",
           "cg_assemble.cu"]
HEADERS = ["cg_types.h", "cg_kernels.cuh", "cg_launch.h", "cg_group.h", "nccl_dl.h", "cg_tma.cuh",
           os.path.join("..", "..", "include", "aphcg.h")]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _nccl_include():
    for p in ("/usr/include/nccl.h",):
        if os.path.exists(p):
            return []
    try:
        import nvidia.nccl  # type: ignore
        for base in nvidia.nccl.__path__:
            inc = os.path.join(base, "include")
            if os.path.exists(os.path.join(inc, "nccl.h")):
                return ["-I" + inc]
    except Exception:
        pass
    return []


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
             "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3"] + _nccl_include()
    nvcc = _nvcc()
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(o)
        procs.append((s, subprocess.Popen([nvcc] + flags + ["-c", os.path.join(CSRC, s), "-o", o],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== %s\n%s\n" % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs
                   + ["-ldl", "-cudart", "static"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
