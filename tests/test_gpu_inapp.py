"""In-app drop-in test: the reference's own application ap.mfer
(built from /root/reference/src into oracle/_ref by oracle/ref/Makefile) runs
examples/202_coalescence (BASELINE config 1: 64^3, walls, two merging bubbles,
density ratio 100) for a few time steps, once with `linsolver_symm = conjugate` and
once with the CUDA module preloaded and selected by name -- nothing else differs.
Every pressure and velocity solve of the projection method then goes through
conjugate_cuda, block decomposition (8 blocks of 32^3) included."""

from __future__ import annotations

import os
import re
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
PLUGIN = os.path.join(ROOT, "aphros_b200", "plugin", "libaphcg_aphros.so")

MESH = "".join("set int %s %d\n" % kv for kv in [
    ("px", 1), ("py", 1), ("pz", 1), ("bx", 2), ("by", 2), ("bz", 2),
    ("bsx", 32), ("bsy", 32), ("bsz", 32)])

LINE = re.compile(r"linear\((\w+)\) '(\w+)': res=(\S+) iter=(\d+)")


MESH_201 = "".join("set int %s %d\n" % kv for kv in [
    ("px", 1), ("py", 1), ("pz", 1), ("bx", 2), ("by", 2), ("bz", 1),
    ("bsx", 16), ("bsy", 16), ("bsz", 1)])


def run_app(tmp, solver, extra, steps, preload_first="", extra_env=None, app="app202", mesh=MESH):
    """preload_first / extra_env: used by tests/test_adapter_cpu.py to put a test double of the
    C ABI in front of libaphcg.so"""
    d = os.path.join(tmp, solver)
    shutil.copytree(os.path.join(REF, app), d)
    with open(os.path.join(d, "mesh.conf"), "w") as f:
        f.write(mesh)
    with open(os.path.join(d, "add.conf"), "w") as f:
        f.write("set string linsolver_symm %s\n" % solver)
        f.write("set string linsolver_gen conjugate\nset string linsolver_vort conjugate\n")
        f.write("set int max_step %d\nset double tmax 100\nset int linreport 1\n" % steps)
        f.write("set int dumppoly 0\nset string dumplist p\nset double dump_field_dt 1e10\n")
        f.write(extra)
    env = dict(os.environ, OMP_NUM_THREADS="8")
    if solver != "conjugate":
        env["LD_PRELOAD"] = ":".join(x for x in (preload_first, PLUGIN, env.get("LD_PRELOAD", "")) if x)
    env.update(extra_env or {})
    p = subprocess.run([os.path.join(REF, "ap.mfer"), "a.conf"], cwd=d, env=env,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    solves = [(m.group(1), m.group(2), float(m.group(3)), int(m.group(4)))
              for m in LINE.finditer(p.stdout + p.stderr)]
    # 202 dumps its fields at the start of the run; 201 only at the end (dumplast)
    dumps = sorted(f for f in os.listdir(d) if re.match(r"p_\d+\.raw$", f))
    pressure = np.fromfile(os.path.join(d, dumps[0 if app == "app202" else -1]), dtype=np.float64)
    with open(os.path.join(d, "stat.dat")) as f:
        head = f.readline().split()
        last = [float(v) for v in f.readlines()[-1].split()]
    return solves, pressure, dict(zip(head, last))


def _need():
    if not (os.path.exists(os.path.join(REF, "ap.mfer")) and os.path.exists(PLUGIN)
            and os.path.isdir(os.path.join(REF, "app202"))):
        pytest.skip("prebuilt ap.mfer / plugin / staged run directory not present")


def test_coalescence_converged_solves(gpu, tmp_path):
    """solves driven to convergence (tol 1e-8 instead of the example's maxiter-limited
    1e-2): the two runs must then agree closely after three full time steps"""
    _need()
    extra = "set double hypre_symm_tol 1e-8\nset int hypre_symm_maxiter 5000\n"
    s_ref, p_ref, st_ref = run_app(str(tmp_path), "conjugate", extra, 3)
    s_gpu, p_gpu, st_gpu = run_app(str(tmp_path), "conjugate_cuda", extra, 3)
    assert len(s_ref) == len(s_gpu) > 10
    assert all(n == "conjugate_cuda" for n, _, _, _ in s_gpu)
    for (_, sys_r, res_r, it_r), (_, sys_g, res_g, it_g) in zip(s_ref, s_gpu):
        assert sys_r == sys_g
        # see tests/cases.py:iterations_ok for the asymmetric bound
        assert it_g <= it_r + 2 + it_r // 100 and it_g >= int(0.97 * it_r) - 2, (sys_r, it_g, it_r)
        assert res_g < 1e-8 or it_g > 5000
    scale = np.abs(p_ref - p_ref.mean()).max()
    assert np.abs((p_gpu - p_gpu.mean()) - (p_ref - p_ref.mean())).max() <= 1e-6 * scale
    for key in ("vol2", "ekin", "pmax"):
        if key in st_ref:
            assert abs(st_gpu[key] - st_ref[key]) <= 1e-6 * max(abs(st_ref[key]), 1e-30), key


def test_coalescence_stock_settings(gpu, tmp_path):
    """the example as shipped (tol 1e-2, miniter 10, maxiter 100): pressure solves hit
    maxiter (101 iterations) in both runs; velocity solves converge in
    the same number of iterations"""
    _need()
    s_ref, p_ref, st_ref = run_app(str(tmp_path), "conjugate", "", 2)
    s_gpu, p_gpu, st_gpu = run_app(str(tmp_path), "conjugate_cuda", "", 2)
    assert len(s_ref) == len(s_gpu)
    for (_, sys_r, res_r, it_r), (_, sys_g, res_g, it_g) in zip(s_ref, s_gpu):
        assert sys_r == sys_g
        if sys_r == "pressure" and it_r == 101:
            assert it_g == 101
        else:
            assert abs(it_g - it_r) <= 2, (sys_r, it_g, it_r)
    # the first time step (zero initial velocity) is identical work: same residuals
    for (_, sys_r, res_r, it_r), (_, sys_g, res_g, it_g) in list(zip(s_ref, s_gpu))[:4]:
        assert abs(res_g - res_r) <= 1e-6 * max(res_r, 1e-30) + 1e-300



def capture_pressure_system(tmp_path, index=6, steps=2):
    """BASELINE config 1 / SURVEY.md 8d input S1: the pressure system of examples/202_coalescence
    (64^3) as the application hands it to the solver, captured by running ap.mfer on the CPU with
    the adapter over the tee-ing test double (tests/cpp/fake_aphcg.c).  index 6 = the first
    pressure solve (after six zero-velocity solves of the start-up).  Returns
    (system (64,64,64,8), x0 (64,64,64), cell_volume)."""
    fake = str(tmp_path / "libfake_aphcg.so")
    subprocess.run(["gcc", "-O2", "-fPIC", "-std=gnu99", "-ffp-contract=off", "-fno-fast-math", "-shared",
                    "-o", fake, os.path.join(ROOT, "tests", "cpp", "fake_aphcg.c"),
                    os.path.join(ROOT, "oracle", "cg_oracle.c"), "-lm"], check=True)
    prefix = str(tmp_path / "tee")
    log = str(tmp_path / "tee_log.txt")
    run_app(str(tmp_path), "conjugate_cuda", "", steps, preload_first=fake,
            extra_env={"FAKE_APHCG_LOG": log, "FAKE_APHCG_TEE": prefix,
                       "FAKE_APHCG_TEE_INDEX": str(index)})
    create = open(log).read().splitlines()[0]
    vol = float(re.search(r"volume=(\S+)", create).group(1))
    system = np.fromfile(prefix + ".sys", dtype=np.float64).reshape(64, 64, 64, 8)
    x0 = np.fromfile(prefix + ".x0", dtype=np.float64).reshape(64, 64, 64)
    return system, x0, vol
