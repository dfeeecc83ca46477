"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/ref_cg =
linear::SolverConjugate / SolverJacobi compiled from /root/reference/src by
oracle/ref/Makefile).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Each fixture stores the inputs' recipe (regenerated from seeds by
aphros_b200.systems, checked by a checksum), the reference's iteration count,
final residual and solution.  Small on purpose (<= 24^3), except

    python tests/golden/make_golden.py assembly   (seconds; needs oracle/_ref/ref_assemble)

which writes assemble_*.npz: rows of the projection pressure system assembled by the
reference's own functions (oracle/ref/ref_assemble.cpp) from a seeded density and seeded face
fluxes -- a SHA-256 per row component (the comparison is bit for bit) plus the first 64 rows; and

    python tests/golden/make_golden.py large      (about 10 minutes, 8 cores, 12 GB)

which adds
  * bench_parity_tlinear64 and bench_parity_64: the two 64^3 systems bench.py cuts into
    z-slabs for its untimed parity check -- the reference unit test's own system (periodic,
    10:1 resistivity, to 1e-12 x the initial residual) and a variable-density system with
    walls (10:1, to 1e-10): iteration count, residual and the whole solution (2 MB each); for
    the second one also the reference's iteration counts for block sizes 8/16/32/64, because
    there its own count moves with the summation order (1726..1732);
  * config5_384_periodic: BASELINE.json config 5 ("384^3 periodic projection solve,
    iterations-to-tolerance vs reference"): the reference's iteration count and residual at
    tol = 1e-7 x initial residual, plus the solution at 4096 sample cells (not the 453 MB field).
"""

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from aphros_b200 import systems  # noqa: E402
from oracle import cpu  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def build_case(name):
    """name -> (system, x0, periodic, kwargs for the solver)"""
    rng = np.random.default_rng(12345)
    if name == "tlinear16_b8":
        s, _ = systems.tlinear_system(16)
        return s, None, (True, True, True), dict(tol=1e-5, maxiter=1000, block=8)
    if name == "tlinear24_b12_guess":
        s, _ = systems.tlinear_system(24)
        return s, rng.standard_normal((24, 24, 24)), (True, True, True), dict(
            tol=1e-7, maxiter=1000, block=12)
    if name == "tlinear_ragged":
        s, _ = systems.tlinear_system(None, shape=(6, 10, 20))
        return s, None, (True, True, True), dict(tol=0.0, maxiter=30, block=(10, 5, 3))
    if name == "density24_neumann":
        s, _ = systems.density_poisson_system(24, nspheres=6, seed=3, rho_in=1e-2)
        tol = 1e-9 * float(np.sqrt((s[..., 7] ** 2).sum() / systems.cell_volume((24, 24, 24))))
        return s, None, (False, False, False), dict(tol=tol, maxiter=3000, block=8)
    if name == "density16_1000to1_fixed":
        s, _ = systems.density_poisson_system(16, nspheres=4, seed=9, rho_in=1e-3)
        return s, None, (False, False, False), dict(tol=0.0, maxiter=100, block=16)
    if name == "const20_periodic_maxnorm":
        s, _ = systems.periodic_constant_system(20)
        return s, None, (True, True, True), dict(tol=1e-4, maxiter=500, block=10, maxnorm=True)
    if name == "tlinear16_miniter":
        s, _ = systems.tlinear_system(16)
        return s, None, (True, True, True), dict(tol=1e30, maxiter=100, miniter=13, block=16)
    if name == "tlinear16_jacobi":
        s, _ = systems.tlinear_system(16)
        return s, None, (True, True, True), dict(tol=1e-4, maxiter=2000, block=8, solver="jacobi")
    raise KeyError(name)


def build_large(name):
    if name == "bench_parity_tlinear64":
        shape = (64, 64, 64)
        s, _ = systems.tlinear_system(64)
        tol = 1e-12 * float(np.sqrt((s[..., 7] ** 2).sum() / systems.cell_volume(shape)))
        return s, None, (True, True, True), dict(tol=tol, maxiter=5000, block=16)
    if name == "bench_parity_64":
        shape = (64, 64, 64)
        s, _ = systems.density_poisson_system(None, nspheres=6, seed=4, rho_in=0.1, shape=shape)
        tol = 1e-10 * float(np.sqrt((s[..., 7] ** 2).sum() / systems.cell_volume(shape)))
        return s, None, (False, False, False), dict(tol=tol, maxiter=5000, block=16)
    if name == "config5_384_periodic":
        shape = (384, 384, 384)
        s, _ = systems.periodic_constant_system(384)
        tol = 1e-7 * float(np.sqrt((s[..., 7] ** 2).sum() / systems.cell_volume(shape)))
        return s, None, (True, True, True), dict(tol=tol, maxiter=5000, block=32)
    raise KeyError(name)


def sample_index(shape, count=4096, seed=777):
    """flat indices of the sample cells of a large fixture"""
    return np.sort(np.random.default_rng(seed).choice(int(np.prod(shape)), count, replace=False))


def main_large():
    assert cpu.have_reference(), "build oracle/_ref first: make -C oracle/ref"
    threads = os.cpu_count() or 1
    workdir = "/dev/shm" if os.path.isdir("/dev/shm") else None
    names = sys.argv[2:] or ["bench_parity_tlinear64", "bench_parity_64", "config5_384_periodic"]
    for name in names:
        s, x0, per, kw = build_large(name)
        x, it, res, sec = cpu.solve_reference(s, x0, periodic=per, threads=threads, workdir=workdir,
                                              **kw)
        out = dict(iter=it, residual=res, tol=kw["tol"], block=kw["block"], threads=threads,
                   seconds=sec)
        if x.size <= 64 ** 3:
            out["x"] = x
            out["system_sha256"] = checksum(s)
            # the reference's own iteration count for other block sizes (summation orders)
            blocks = [8, 16, 32, 64]
            out["blocks"] = np.array(blocks)
            out["iter_by_block"] = np.array([
                it if b == kw["block"] else cpu.solve_reference(
                    s, x0, periodic=per, threads=threads, workdir=workdir, **dict(kw, block=b))[1]
                for b in blocks])
            print(name, "iterations by block size", dict(zip(blocks, out["iter_by_block"].tolist())),
                  flush=True)
        else:
            idx = sample_index(x.shape)
            out["sample_index"] = idx
            out["sample_x"] = x.reshape(-1)[idx]
            out["x_mean"] = float(x.mean())
            out["rhs_sha256"] = checksum(s[..., 7])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, it, res, "%.1f s" % sec, flush=True)


ASSEMBLY_CASES = ["assemble_walls32", "assemble_perz32_b16", "assemble_perxz_ragged"]


def build_assembly_case(name):
    """name -> dict(rho, vx, vy, vz, source, dt, periodic, block): inputs of the projection
    assembly (density spanning 4 orders of magnitude, random face fluxes)"""
    shape, per, block = {
        "assemble_walls32": ((32, 32, 32), (False, False, False), None),
        "assemble_perz32_b16": ((32, 32, 32), (False, False, True), 16),
        "assemble_perxz_ragged": ((12, 10, 16), (True, False, True), None),
    }[name]
    nz, ny, nx = shape
    rng = np.random.default_rng(20240700 + len(name))
    rho = np.exp(rng.standard_normal(shape) * 2)
    vx = rng.standard_normal((nz, ny, nx + 1))
    vy = rng.standard_normal((nz, ny + 1, nx))
    vz = rng.standard_normal((nz + 1, ny, nx))
    if per[0]:
        vx[:, :, -1] = vx[:, :, 0]
    if per[1]:
        vy[:, -1, :] = vy[:, 0, :]
    if per[2]:
        vz[-1] = vz[0]
    src = rng.standard_normal(shape)
    return dict(rho=rho, vx=vx, vy=vy, vz=vz, source=src, dt=1e-3, periodic=per, block=block)


def consistent_projection_case(name):
    """like build_assembly_case, but a system CG can solve: moderate density variation, no
    source, zero flux through walls -- the right-hand side then sums to zero (telescoping), as
    the singular Neumann / periodic problem requires"""
    c = build_assembly_case(name)
    rng = np.random.default_rng(99)
    c["rho"] = np.exp(0.5 * rng.standard_normal(c["rho"].shape))
    per = c["periodic"]
    if not per[0]:
        c["vx"][:, :, 0] = c["vx"][:, :, -1] = 0.0
    if not per[1]:
        c["vy"][:, 0, :] = c["vy"][:, -1, :] = 0.0
    if not per[2]:
        c["vz"][0] = c["vz"][-1] = 0.0
    c["source"] = None
    return c


def main_assembly():
    assert cpu.have_reference_assembler(), "build oracle/_ref first: make -C oracle/ref app"
    for name in ASSEMBLY_CASES:
        c = build_assembly_case(name)
        rows = cpu.assemble_reference(c["rho"], c["vx"], c["vy"], c["vz"], c["source"], dt=c["dt"],
                                      periodic=c["periodic"], block=c["block"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            # "+ 0.0": a zero wall coefficient may be -0.0 or +0.0
                            sha256=np.array([checksum(rows[..., q] + 0.0) for q in range(8)]),
                            first_rows=rows.reshape(-1, 8)[:64],
                            inputs_sha256=checksum(np.concatenate(
                                [c[k].ravel() for k in ("rho", "vx", "vy", "vz", "source")])))
        print(name, rows.shape, float(np.abs(rows[..., 0]).max()))


CASES = ["tlinear16_b8", "tlinear24_b12_guess", "tlinear_ragged", "density24_neumann",
         "density16_1000to1_fixed", "const20_periodic_maxnorm", "tlinear16_miniter",
         "tlinear16_jacobi"]


def checksum(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert cpu.have_reference(), "build oracle/_ref first: make -C oracle/ref"
    for name in CASES:
        s, x0, per, kw = build_case(name)
        x, it, res, _ = cpu.solve_reference(s, x0, periodic=per, **kw)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, iter=it, residual=res,
                            system_sha256=checksum(s))
        print(name, it, res)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "large":
        main_large()
    elif len(sys.argv) > 1 and sys.argv[1] == "assembly":
        main_assembly()
    else:
        main()
