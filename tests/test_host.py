"""CPU tests of the host side: the C ABI library loads and exports every symbol
include/aphcg.h declares (no compute without a GPU), the interface mirror keeps
the reference's semantics, slab logic, and a world_size-2 gloo run of the
multi-rank plumbing."""

from __future__ import annotations

import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from aphros_b200 import Conf, Mesh, ModuleLinear, SolverConjugateCuda, capi, distr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "aphcg.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(aphcg_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libaphcg.so does not export " + n
    assert set(names) == set(capi.SIGNATURES), set(names) ^ set(capi.SIGNATURES)
    assert capi.lib().aphcg_version() == 1


def test_sass_is_blackwell_native(built):
    """the shipped kernels carry TMA tensor loads and 128-bit global accesses"""
    out = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    assert "UTMALDG.3D" in out.stdout
    assert "LDG.E.EF.128" in out.stdout or "LDG.E.128" in out.stdout
    assert "SYNCS.ARRIVE.TRANS64" in out.stdout  # mbarrier expect_tx


@pytest.mark.skipif(capi.lib().aphcg_device_count() > 0, reason="GPU present")
def test_fails_loudly_without_gpu(built):
    with pytest.raises(capi.AphcgError, match="no CUDA device"):
        SolverConjugateCuda(Conf(), {}, Mesh(shape=(8, 8, 8)))


def test_group_argument_checks_and_partition(built):
    """aphcg_group_*: argument checks run before any CUDA call; without a GPU creation
    fails loudly; the slab rule is distr.slab_partition's"""
    from aphros_b200 import SolverConjugateCudaGroup
    with pytest.raises(capi.AphcgError, match="cannot cut"):
        SolverConjugateCudaGroup(Conf(), {}, Mesh(shape=(3, 8, 8)), [0, 0, 0, 0])
    with pytest.raises(capi.AphcgError, match="1..16 slabs"):
        SolverConjugateCudaGroup(Conf(), {}, Mesh(shape=(64, 8, 8)), [0] * 17)
    if capi.lib().aphcg_device_count() == 0:
        with pytest.raises(capi.AphcgError, match="no CUDA device"):
            SolverConjugateCudaGroup(Conf(), {}, Mesh(shape=(8, 8, 8)), [0, 1])
    mod = ModuleLinear.GetInstance("conjugate_cuda")
    var = {"hypre_symm_tol": 1e-3, "hypre_symm_maxiter": 10, "cuda_devices": 20}
    with pytest.raises(capi.AphcgError, match="1..16 slabs"):
        mod.Make(var, "symm", Mesh(shape=(64, 8, 8)))
    var = {"hypre_symm_tol": 1e-3, "hypre_symm_maxiter": 10, "cuda_devices": 3,
           "cuda_slabs_per_device": 6}
    with pytest.raises(capi.AphcgError, match="1..16 slabs"):
        mod.Make(var, "symm", Mesh(shape=(64, 8, 8)))


@pytest.mark.parametrize("world", [1, 2])
def test_bench_line_contract_with_stubbed_solver(world):
    """bench.py's JSON line (the driver's contract) assembled with the CUDA solver, torch.cuda
    and torch.distributed stubbed out: every required key, at N = 1 and N > 1"""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "helpers", "bench_stub.py"),
                          str(world)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "clocks",
                "e2e", "gpu_launches", "roofline", "residual"):
        assert key in line, key
    assert line["n_gpus"] == world and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e_projection"])
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert {"rel_max_abs", "iter", "iter_reference", "ok"} <= set(line["parity"])
    assert ("strong" in line) == (world > 1)
    if world > 1:
        assert {"value", "n1_value", "efficiency_vs_n1"} <= set(line["strong"])
    if world == 1:
        assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
        assert line["roofline"]["traffic"] and line["roofline"]["dram"]["frac"] < 1.0
        assert line["roofline"]["frac_dram"] == line["roofline"]["dram"]["frac"]
        assert abs(line["roofline"]["frac"] - 120.0 * 512 ** 3 / 1.89e-3 / 1e9 / line["roofline"]["peak"]) < 1e-9
    else:
        assert "cpu_baseline" not in line


def test_group_sync_barrier(tmp_path):
    """the thread barrier / host all-reduce of the in-process slab group (cg_group.h)"""
    exe = str(tmp_path / "group_sync_test")
    subprocess.run(["g++", "-std=c++17", "-O1", "-pthread", "-I" + os.path.join(ROOT, "aphros_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "group_sync_test.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def test_argument_checks(built):
    L = capi.lib()
    assert L.aphcg_group_size(None) == 0 and L.aphcg_stream(None) is None
    assert L.aphcg_group_member(None, 0) is None and L.aphcg_launch_count(None) == 0
    h = ctypes.c_void_p()
    d = capi.Desc()
    d.nx, d.ny, d.nz = 0, 4, 4
    assert L.aphcg_create(ctypes.byref(h), ctypes.byref(d)) == -1
    assert b"bad mesh size" in L.aphcg_last_error()
    d.nx = 4
    d.nranks, d.rank, d.nz_local, d.cell_volume = 2, 2, 2, 1.0
    assert L.aphcg_create(ctypes.byref(h), ctypes.byref(d)) == -1
    assert L.aphcg_run(None, None, None) == -1
    assert L.aphcg_destroy(None) == 0


def test_module_registry_and_conf_keys():
    """ModuleLinear<M>::GetConf reads hypre_<prefix>_{tol,maxiter,miniter}
    (src/linear/linear.h:66-75): tol/maxiter mandatory, miniter defaults to 0"""
    assert set(ModuleLinear.GetInstances()) >= {"conjugate_cuda", "jacobi_cuda"}
    assert ModuleLinear.GetInstance("nope") is None
    var = {"hypre_symm_tol": 1e-3, "hypre_symm_maxiter": 100}
    c = ModuleLinear.GetConf(var, "symm")
    assert (c.tol, c.maxiter, c.miniter) == (1e-3, 100, 0)
    with pytest.raises(KeyError):
        ModuleLinear.GetConf({"hypre_symm_tol": 1.0}, "symm")
    with pytest.raises(RuntimeError, match="already registered"):
        ModuleLinear.Register(ModuleLinear.GetInstance("conjugate_cuda"))
    assert Conf().miniter == 1 and Conf().maxiter == 100 and Conf().tol == 0  # linear.h:21-25


def test_layout_of_reference_field():
    """a FieldCell with hl=2 halos and one padding cell (src/geom/mesh.ipp:60-113)"""
    n, hl = 8, 2
    full = n + 2 * hl + 1
    a = np.zeros((full, full, full))
    v = a[hl:hl + n, hl:hl + n, hl:hl + n]
    lay = capi.layout_of(v, (n, n, n))
    assert (lay.offset, lay.stride_y, lay.stride_z) == (0, full, full * full)
    rows = np.zeros((full, full, full, 8))
    lay = capi.layout_of(rows[hl:hl + n, hl:hl + n, hl:hl + n], (n, n, n), 8)
    assert (lay.stride_y, lay.stride_z) == (full, full * full)
    with pytest.raises(ValueError):
        capi.layout_of(a[:, :, ::2], (full, full, full // 2))


def test_slab_partition():
    assert distr.slab_partition(512, 8) == [(64 * r, 64) for r in range(8)]
    parts = distr.slab_partition(10, 4)
    assert parts == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert sum(n for _, n in parts) == 10
    with pytest.raises(ValueError):
        distr.slab_partition(3, 4)
    assert distr.neighbours(0, 4, False) == (None, 1)
    assert distr.neighbours(3, 4, False) == (2, None)
    assert distr.neighbours(0, 4, True) == (3, 1)
    assert distr.neighbours(1, 2, True) == (0, 0)
    m = distr.local_mesh((10, 6, 4), (True, True, False), 2, 4)
    assert (m.z0, m.nz_local, m.local_shape, m.device) == (6, 2, (2, 6, 4), 2)


def _bench_shape():
    sys.path.insert(0, ROOT)
    import bench
    return bench.global_shape


def test_bench_weak_scaling_shapes():
    gs = _bench_shape()
    assert gs(1, 512) == (512, 512, 512)
    assert gs(2, 512) == (1024, 512, 512)
    assert gs(4, 512) == (1024, 1024, 512)
    assert gs(8, 512) == (1024, 1024, 1024)


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from aphros_b200 import distr, systems
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
# 1. byte plumbing used for the NCCL id and the IPC handles
blob = bytes([rank + 1] * 128)
got = distr.all_gather_bytes(blob)
assert got == [bytes([r + 1] * 128) for r in range(world)], got
uid = distr.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128, 0)
assert uid == bytes(range(128))
# 2. slab decomposition: every rank takes its planes of the same global system;
#    with the neighbours' boundary planes (what the kernels write into the ghost
#    planes) each slab reproduces the global operator exactly
shape = (12, 6, 8)
per = (True, True, True)
s, _ = systems.tlinear_system(None, shape=shape)
v = np.random.default_rng(1).standard_normal(shape)
m = distr.local_mesh(shape, per, rank, world)
lo, hi = distr.neighbours(rank, world, True)
sl = slice(m.z0, m.z0 + m.nz_local)
planes = [None] * world
dist.all_gather_object(planes, (v[sl][0].copy(), v[sl][-1].copy()))
ghost_lo, ghost_hi = planes[lo][1], planes[hi][0]
vv = np.concatenate([ghost_lo[None], v[sl], ghost_hi[None]])
a = s[sl]
out = vv[1:-1] * a[..., 0]
out = out + np.roll(vv[1:-1], 1, axis=2) * a[..., 1] + np.roll(vv[1:-1], -1, axis=2) * a[..., 2]
out = out + np.roll(vv[1:-1], 1, axis=1) * a[..., 3] + np.roll(vv[1:-1], -1, axis=1) * a[..., 4]
out = out + vv[:-2] * a[..., 5] + vv[2:] * a[..., 6]
from oracle import cpu
ref = cpu.apply(s, v, periodic=per)[sl]
assert np.allclose(out, ref, rtol=0, atol=1e-12 * np.abs(ref).max()), np.abs(out - ref).max()
parts = [None] * world
dist.all_gather_object(parts, (m.z0, m.nz_local))
assert parts == distr.slab_partition(shape[0], world)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_plumbing_gloo(built, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", WORLD_SIZE="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o)
        assert "rank %d ok" % r in o


def test_bench_host_helpers():
    """bench.py's host-side guards: the memory figure honours cgroup limits, the scratch
    directory has room, NUMA binding is a no-op where the topology cannot be read"""
    sys.path.insert(0, ROOT)
    import bench
    from aphros_b200 import distr
    gb = bench.host_ram_gb()
    assert 0.0 < gb < 1e5
    assert bench.scratch_dir(1) is not None
    assert bench.scratch_dir(10 ** 18) is None
    before = os.sched_getaffinity(0)
    assert distr.bind_to_gpu_numa_node(0) is None or isinstance(distr.bind_to_gpu_numa_node(0), str)
    os.sched_setaffinity(0, before)
