"""CPU test of the aphros adapter's HOST logic (aphros_b200/plugin/linear_conjugate_cuda.cpp).

The reference's own classes (oracle/_ref/ref_cg: DistrSolver, MeshCartesian, stage coroutines,
the ModuleLinear factory) load the adapter and select `conjugate_cuda` by name, exactly as in
tests/test_gpu_dropin.py -- but the C ABI underneath is answered by a TEST DOUBLE
(tests/cpp/fake_aphcg.c, preloaded in front of libaphcg.so) that forwards to the CPU oracle and
logs every call.  What is checked here is what the adapter does around the C ABI: gather of the
blocks' rows into rank-wide arrays, scatter of the solution and its halo exchange, the geometry,
periodicity, cell volume, flags, Conf and device list it passes.  The product has no CPU path;
the arithmetic of the CUDA library is tested by the `-m gpu` tests."""

from __future__ import annotations

import os
import re
import subprocess
import sys

import numpy as np
import pytest

from aphros_b200 import systems
from cases import rel_max_abs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "aphros_b200", "plugin", "libaphcg_aphros.so")


@pytest.fixture(scope="module")
def fake(built, tmp_path_factory):
    from oracle import cpu
    if not (cpu.have_reference() and os.path.exists(PLUGIN)):
        pytest.skip("prebuilt oracle/_ref and adapter not present")
    d = tmp_path_factory.mktemp("fake")
    so = str(d / "libfake_aphcg.so")
    subprocess.run(["gcc", "-O2", "-fPIC", "-std=gnu99", "-ffp-contract=off", "-fno-fast-math", "-shared",
                    "-o", so, os.path.join(ROOT, "tests", "cpp", "fake_aphcg.c"),
                    os.path.join(ROOT, "oracle", "cg_oracle.c"), "-lm"], check=True)
    return so


def run(fake, tmp_path, system, x0=None, solver="conjugate_cuda", **kw):
    """(x, iter, residual, log lines) of one ref_cg run with the adapter over the test double"""
    from oracle import cpu
    log = str(tmp_path / ("log_%d.txt" % len(os.listdir(tmp_path))))
    env = {"LD_PRELOAD": fake, "FAKE_APHCG_LOG": log}
    x, it, res, _ = cpu.solve_reference(system, x0, solver=solver, plugin=PLUGIN, env=env, **kw)
    with open(log) as f:
        lines = f.read().splitlines()
    return x, it, res, lines


@pytest.mark.parametrize("backend", ["native", "local"])
@pytest.mark.parametrize("block,threads", [(16, 1), (8, 4), (32, 1)])
def test_gather_scatter_over_blocks(fake, tmp_path, backend, block, threads):
    """8 / 64 / 1 blocks per rank, serial and OpenMP over blocks: the rank-wide system the
    adapter assembles is the reference's, so the answer is `conjugate`'s"""
    from oracle import cpu
    s, _ = systems.tlinear_system(32)
    kw = dict(tol=1e-9, maxiter=2000, block=block, threads=threads,
              extra="set string backend %s" % backend)
    xr, itr, resr, _ = cpu.solve_reference(s, solver="conjugate", **kw)
    xg, itg, resg, log = run(fake, tmp_path, s, **kw)
    assert abs(itg - itr) <= 2 and resg < 1e-9
    assert rel_max_abs(xg, xr) <= 1e-10
    create = [l for l in log if l.startswith("create")]
    assert len(create) == 1, "one device object per rank, owned by the lead block"
    m = re.match(r"create nx=32 ny=32 nz=32 periodic=111 volume=(\S+) flags=0 devices=\[0\]", create[0])
    assert m, create[0]
    assert abs(float(m.group(1)) - cpu.reference_cell_volume((32, 32, 32), block)) < 1e-20
    assert sum(l.startswith("solve") for l in log) == 1 and log[-1] == "destroy"
    # (the driver always passes an fc_init field, zero when no guess is given)
    assert "tol=1.0000000000000001e-09 miniter=0 maxiter=2000" in [l for l in log if l.startswith("solve")][0]


@pytest.mark.parametrize("name,threads", [("assemble_walls32", 1), ("assemble_perz32_b16", 4),
                                          ("assemble_perxz_ragged", 1)])
def test_projection_entry_gathers_density_and_fluxes(fake, tmp_path, name, threads):
    """linear::ProjectionSolver (aphros_b200/plugin/linear_projection.h) over the test double:
    the adapter gathers the blocks' density (with its z ghost planes), face fluxes and source
    into the rank-wide arrays of aphcg_group_assemble_projection; the double assembles rows
    from exactly those arrays, and the solve must be the reference's own for the rows the
    reference's assembler builds from the same fields (one block, 8 blocks with OpenMP,
    periodic x/z on a ragged mesh)"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    from oracle import cpu
    if not cpu.have_reference_assembler():
        pytest.skip("oracle/_ref/ref_assemble not present")
    c = make_golden.consistent_projection_case(name)
    log = str(tmp_path / "log_proj.txt")
    kw = dict(tol=1e-9, maxiter=3000)
    rows, x, it, res = cpu.assemble_reference(
        c["rho"], c["vx"], c["vy"], c["vz"], c["source"], dt=c["dt"], periodic=c["periodic"],
        block=c["block"], plugin=PLUGIN, threads=threads,
        env={"LD_PRELOAD": fake, "FAKE_APHCG_LOG": log}, **kw)
    xr, itr, resr, _ = cpu.solve_reference(rows, periodic=c["periodic"], block=c["block"],
                                           solver="conjugate", **kw)
    assert itr < 3000 and abs(it - itr) <= 2 + itr // 100 and res < 1e-9
    assert rel_max_abs(x, xr) <= 1e-7
    lines = open(log).read().splitlines()
    assert any(l.startswith("assemble_projection dt=0.001") and "source=0" in l for l in lines)
    assert [l.split()[0] for l in lines if not l.startswith("create")] == [
        "assemble_projection", "upload_guess", "run", "download_solution", "destroy"]


def test_guess_walls_maxnorm_and_ragged_mesh(fake, tmp_path):
    """non-cubic mesh, Neumann walls (periodic flags 000), an initial guess (fc_init != nullptr),
    `linsolver_symm_maxnorm` -> APHCG_MAXNORM"""
    from oracle import cpu
    shape = (16, 24, 32)
    s, _ = systems.density_poisson_system(None, nspheres=4, seed=3, rho_in=0.2, shape=shape)
    x0 = np.random.default_rng(5).standard_normal(shape) * 1e-3
    kw = dict(periodic=(False, False, False), tol=0.0, maxiter=25, block=8, maxnorm=True)
    xr, itr, resr, _ = cpu.solve_reference(s, x0, solver="conjugate", **kw)
    xg, itg, resg, log = run(fake, tmp_path, s, x0, **kw)
    assert itg == itr == 26
    # the test double sums in one-block order, the reference per 8^3 block: rounding-level
    # differences, amplified by the 5:1 density jump over 26 non-converged iterations
    assert abs(resg - resr) <= 1e-7 * resr
    assert rel_max_abs(xg, xr) <= 1e-7
    create = [l for l in log if l.startswith("create")][0]
    assert "nx=32 ny=24 nz=16 periodic=000" in create and "flags=1 " in create
    assert "solve guess=1" in "\n".join(log)


def test_config_keys_reach_the_c_abi(fake, tmp_path):
    """cuda_device / cuda_devices / cuda_slabs_per_device -> device list; linsolver_symm_cuda_graph,
    _cuda_tma, _jacobi -> flags (APHCG_NO_GRAPH=2, APHCG_NO_TMA=4, APHCG_JACOBI_PRECOND=32)"""
    s, _ = systems.tlinear_system(16)
    kw = dict(tol=1e-3, maxiter=50, block=8)
    cases = [("set int cuda_device 2\nset int cuda_devices 3", "devices=[2,3,4]", "flags=0 "),
             ("set int cuda_slabs_per_device 2", "devices=[0,0]", "flags=0 "),
             ("set int cuda_devices 2\nset int cuda_slabs_per_device 2", "devices=[0,0,1,1]", "flags=0 "),
             ("set int linsolver_symm_cuda_graph 0", "devices=[0]", "flags=2 "),
             ("set int linsolver_symm_cuda_tma 0\nset int linsolver_symm_jacobi 1", "devices=[0]", "flags=36 ")]
    for extra, dev, flags in cases:
        _, _, _, log = run(fake, tmp_path, s, extra=extra, **kw)
        create = [l for l in log if l.startswith("create")][0]
        assert dev in create and flags in create, (extra, create)
    with pytest.raises(RuntimeError, match="cuda_devices x cuda_slabs_per_device must be 1..16"):
        run(fake, tmp_path, s, extra="set int cuda_devices 17", **kw)


def test_jacobi_module_calls(fake, tmp_path):
    """jacobi_cuda: upload_system, upload_guess, run_jacobi, download_solution, same answer as
    the reference's `jacobi`"""
    from oracle import cpu
    s, _ = systems.tlinear_system(16)
    kw = dict(tol=1e-4, maxiter=2000, block=8)
    xr, itr, resr, _ = cpu.solve_reference(s, solver="jacobi", **kw)
    xg, itg, resg, log = run(fake, tmp_path, s, solver="jacobi_cuda", **kw)
    assert itg == itr and rel_max_abs(xg, xr) <= 1e-12
    calls = [l.split()[0] for l in log]
    assert calls == ["create", "upload_system", "upload_guess", "run_jacobi", "download_solution",
                     "destroy"], calls


def test_repeated_solves_reuse_the_device_object(fake, tmp_path):
    """SetConf / Solve every step on the same solver objects (src/kernel/hydro.ipp:2351-2353):
    one create, N solves"""
    s, _ = systems.tlinear_system(16)
    _, _, _, log = run(fake, tmp_path, s, tol=1e-6, maxiter=500, block=8, repeat=3)
    assert sum(l.startswith("create") for l in log) == 1
    assert sum(l.startswith("solve") for l in log) == 3


PLUGIN2 = os.path.join(ROOT, "aphros_b200", "plugin", "libaphcg_aphros2.so")


def system_2d(shape2, walls):
    """a 2-D 5-point system in the 8-double row format: the generators' (1, ny, nx) system
    with its z faces removed (their coefficients folded out of the diagonal)"""
    ny, nx = shape2
    if walls:
        s, _ = systems.density_poisson_system(None, nspheres=3, seed=2, rho_in=0.2, shape=(1, ny, nx))
    else:
        s, _ = systems.tlinear_system(None, shape=(1, ny, nx))
    s = s.copy()
    s[..., 0] += s[..., 5] + s[..., 6]
    s[..., 5:7] = 0.0
    return s


@pytest.mark.parametrize("walls", [False, True])
def test_two_dimensional_mesh(fake, tmp_path, walls):
    """MeshCartesian<double,2> (the reference registers `conjugate` for every enabled dimension,
    src/linear/linear.cpp:16-23): 6-double rows widened to the 8-double format, nz = 1, the
    missing direction neither periodic nor coupled; same answer as the reference's 2-D
    `conjugate` over 4x4 blocks"""
    from oracle import cpu
    if not (cpu.have_reference_dim2() and os.path.exists(PLUGIN2)):
        pytest.skip("2-D reference build (make -C oracle/ref dim2) not present")
    s = system_2d((48, 64), walls)
    per = (not walls, not walls, False)
    kw = dict(periodic=per, tol=1e-9, maxiter=3000, block=(16, 12, 1), dim=2)
    xr, itr, resr, _ = cpu.solve_reference(s, solver="conjugate", **kw)
    log = str(tmp_path / "log2d.txt")
    xg, itg, resg, _ = cpu.solve_reference(s, solver="conjugate_cuda", plugin=PLUGIN2,
                                           env={"LD_PRELOAD": fake, "FAKE_APHCG_LOG": log}, **kw)
    assert abs(itg - itr) <= 2 and resg < 1e-9
    assert rel_max_abs(xg, xr) <= (1e-8 if walls else 1e-10)
    create = [l for l in open(log).read().splitlines() if l.startswith("create")][0]
    assert "nx=64 ny=48 nz=1 periodic=%d%d0" % (per[0], per[1]) in create
    # cell "volume" of a 2-D mesh is the cell area (m.GetCellSize().prod(), mesh.h:171-173)
    vol = float(re.search(r"volume=(\S+)", create).group(1))
    assert abs(vol - (1.0 / 64) ** 2) < 1e-12


def test_inapp_coalescence_over_the_test_double(fake, tmp_path):
    """The reference's whole application (oracle/_ref/ap.mfer, examples/202_coalescence, 64^3,
    8 blocks of 32^3, 8 OpenMP threads) with the adapter preloaded and selected by
    `linsolver_symm = conjugate_cuda`, the C ABI answered by the test double: ONE device object
    per rank serves the pressure solve and the three velocity solves of every step
    (src/kernel/hydro.ipp:698-716), SetConf values arrive (tol 1e-2, miniter 10, maxiter 100),
    and the solve sequence is that of the stock `conjugate` run."""
    import test_gpu_inapp as app
    if not (os.path.exists(os.path.join(app.REF, "ap.mfer")) and os.path.isdir(os.path.join(app.REF, "app202"))):
        pytest.skip("prebuilt ap.mfer / staged run directory not present")
    log = str(tmp_path / "log_app.txt")
    s_ref, p_ref, st_ref = app.run_app(str(tmp_path), "conjugate", "", 2)
    s_dbl, p_dbl, st_dbl = app.run_app(str(tmp_path), "conjugate_cuda", "", 2, preload_first=fake,
                                       extra_env={"FAKE_APHCG_LOG": log})
    lines = open(log).read().splitlines()
    assert sum(l.startswith("create") for l in lines) == 1
    assert "nx=64 ny=64 nz=64 periodic=000" in lines[0] and lines[-1] == "destroy"
    solves = [l for l in lines if l.startswith("solve")]
    assert len(solves) == len(s_ref) == len(s_dbl) > 8
    assert all("guess=1 tol=0.01 miniter=10 maxiter=100" in l for l in solves)
    for (_, sys_r, res_r, it_r), (name, sys_d, res_d, it_d) in zip(s_ref, s_dbl):
        assert name == "conjugate_cuda" and sys_r == sys_d
        assert abs(it_d - it_r) <= 2, (sys_r, it_d, it_r)
    # known answers of the first step: pressure hits maxiter, velocity 28/26/26
    its = [it for _, _, _, it in s_dbl]
    assert 101 in its and {26, 28} <= set(its)
    scale = np.abs(p_ref - p_ref.mean()).max()
    assert np.abs((p_dbl - p_dbl.mean()) - (p_ref - p_ref.mean())).max() <= 1e-3 * scale


def test_inapp_taylor_couette_over_the_test_double(fake, tmp_path):
    """examples/201_taylor_couette through ap.mfer with the adapter over the test double:
    a 32x32x1 mesh (`dim 2`, periodic in z) with embedded boundaries -- identity rows for
    excluded cells, a zero system for the third velocity component -- in 4 blocks of 16x16x1"""
    import test_gpu_inapp as app
    if not (os.path.exists(os.path.join(app.REF, "ap.mfer")) and os.path.isdir(os.path.join(app.REF, "app201"))):
        pytest.skip("prebuilt ap.mfer / staged run directory not present")
    log = str(tmp_path / "log_app201.txt")
    extra = "set int hypre_symm_maxiter 1000\n"
    kw = dict(app="app201", mesh=app.MESH_201)
    s_ref, p_ref, _ = app.run_app(str(tmp_path), "conjugate", extra, 3, **kw)
    s_dbl, p_dbl, _ = app.run_app(str(tmp_path), "conjugate_cuda", extra, 3, preload_first=fake,
                                  extra_env={"FAKE_APHCG_LOG": log}, **kw)
    lines = open(log).read().splitlines()
    assert sum(l.startswith("create") for l in lines) == 1
    assert "nx=32 ny=32 nz=1 periodic=001" in lines[0]
    assert len(s_ref) == len(s_dbl) >= 12
    for (_, sys_r, res_r, it_r), (name, sys_d, res_d, it_d) in zip(s_ref, s_dbl):
        assert name == "conjugate_cuda" and sys_r == sys_d
        assert abs(it_d - it_r) <= 2 and res_d < 1e-7, (sys_r, it_d, it_r, res_d)
    scale = np.abs(p_ref - p_ref.mean()).max()
    assert np.abs((p_dbl - p_dbl.mean()) - (p_ref - p_ref.mean())).max() <= 1e-6 * scale


def test_capture_of_a_live_system(fake, tmp_path):
    """`linsolver_symm_cuda_dump PREFIX` (+ `_cuda_dump_index N`): the adapter writes the
    rank-wide rows and guess of the N-th Solve exactly as they go to the C ABI, with a text
    header -- the raw stand-in for the reference's HDF5 `--system_out`"""
    from aphros_b200 import tlinear
    shape = (16, 24, 32)
    s, _ = systems.density_poisson_system(None, nspheres=4, seed=3, rho_in=0.2, shape=shape)
    x0 = np.random.default_rng(5).standard_normal(shape) * 1e-3
    prefix = str(tmp_path / "cap")
    extra = "set string linsolver_symm_cuda_dump %s\nset int linsolver_symm_cuda_dump_index 1" % prefix
    run(fake, tmp_path, s, x0, periodic=(False, True, False), tol=1e-5, maxiter=77, miniter=3,
        block=8, repeat=2, extra=extra)
    meta, system, guess = tlinear.read_capture(prefix)
    assert meta["shape"] == shape and meta["periodic"] == (False, True, False)
    assert (meta["tol"], meta["miniter"], meta["maxiter"], meta["name"]) == (1e-5, 3, 77, "pressure")
    assert np.array_equal(system, s) and np.array_equal(guess, x0)
    from oracle import cpu
    assert abs(meta["cell_volume"] - cpu.reference_cell_volume(shape, 8)) < 1e-20
    assert "call 1" in open(prefix + ".txt").read()


def test_captured_config1_system_reproduces_the_reference_log(fake, tmp_path):
    """BASELINE config 1 on the CPU side: the 64^3 pressure system of example 202 captured from
    the application (tee of the test double), solved by the oracle with the example's settings,
    gives the reference's own log line `res=1.40709591e+00 iter=101` (8 blocks of 32^3); the
    system is bitwise symmetric with zero wall coefficients, i.e. 4-stream storage applies"""
    import test_gpu_inapp as app
    from oracle import cpu
    if not (os.path.exists(os.path.join(app.REF, "ap.mfer")) and os.path.isdir(os.path.join(app.REF, "app202"))):
        pytest.skip("prebuilt ap.mfer / staged run directory not present")
    s, x0, vol = app.capture_pressure_system(tmp_path)
    assert vol == 0.001953125 and not x0.any()
    _, it, res, _ = cpu.solve(s, x0, periodic=(False, False, False), cell_volume=vol, tol=1e-2,
                              miniter=10, maxiter=100, block=32)
    assert it == 101 and abs(res - 1.40709591) < 5e-9
    assert np.array_equal(s[:, :, :-1, 2], s[:, :, 1:, 1])
    assert np.array_equal(s[:, :-1, :, 4], s[:, 1:, :, 3])
    assert np.array_equal(s[:-1, :, :, 6], s[1:, :, :, 5])
    assert not s[:, :, 0, 1].any() and not s[:, :, -1, 2].any() and not s[0, :, :, 5].any()
