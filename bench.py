#!/usr/bin/env python
"""Benchmark of the pressure-Poisson CG hot path (BASELINE.json metric:
"Poisson CG cell-iter/s at 512^3 FP64, 1/2/4/8 B200; % of HBM roofline").

    python bench.py --gpus N --steps K --warmup W            (N=1)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                     (reference CPU CG)

A "step" is one solve of the reference's own benchmark setting
(src/test/linear/run_bench: --tol 0 --maxiter 100 -> 101 CG iterations) on a
synthetic variable-density projection system (SURVEY.md 8(d) S3/S4, 1000:1
density jump, Neumann walls) with 512^3 cells per GPU; weak scaling: the domain
doubles in z, y, x as N doubles (N=8: 1024^3) and is cut into z-slabs.
Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "poisson_cg_cell_iterations_per_second"
UNIT = "cell-iter/s"
B_ALG = 144.0            # algorithmic bytes per cell-iteration
B_ALG_DIR_SPMV = 120.0   # share of the fused direction+SpMV kernel
B_ALG_UPDATE = 24.0      # share of the update kernel
MAXITER = 100            # -> 101 iterations per step
SEEDS = {1: (512, 20240602), 2: (1024, 20240604), 4: (2048, 20240605), 8: (4096, 20240603)}


def global_shape(n_gpus: int, per_gpu: int):
    """(nz, ny, nx): per_gpu^3 cells per GPU, doubling z, y, x in turn."""
    dims = [per_gpu, per_gpu, per_gpu]
    k, i = n_gpus, 0
    while k > 1:
        dims[i % 3] *= 2
        k //= 2
        i += 1
    return tuple(dims)


def measured_traffic(cells, kernel_tag):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture
    (profiles/r01d_traffic.json), if it was taken on this workload; else None."""
    p = os.path.join(ROOT, "profiles", "r01d_traffic.json")
    try:
        with open(p) as f:
            t = json.load(f)
        if t["cells"] == cells and kernel_tag in t["kernels"]:
            return t["kernels"][kernel_tag]["k_dir_spmv_bytes_per_launch"]
    except Exception:
        pass
    return None


def peak_hbm_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own SolverConjugate (oracle/_ref)
# ---------------------------------------------------------------------------------
def run_reference_cpu(steps: int, warmup: int, budget_s: float = 30.0):
    """Times the reference's CPU CG (built from its own sources into oracle/_ref,
    OpenMP over 32^3 blocks) on a bounded sample of the bench workload."""
    from aphros_b200 import systems
    from oracle import cpu  # the one place bench.py may execute oracle/

    cores = os.cpu_count() or 1
    n = 256 if cores >= 16 else 128
    nsph = 64 if n == 256 else 8
    system, _ = systems.density_poisson_system(n, nspheres=nsph, seed=20240601)
    kind = "reference" if cpu.have_reference() else "port"
    sample = ("%d^3 S2 variable-density system (1000:1), tol=0 maxiter=%d (%d iterations), "
              % (n, MAXITER, MAXITER + 1))
    if kind == "reference":
        workdir = "/dev/shm" if os.path.isdir("/dev/shm") else None
        # one process, `repeat` solves: best-of time of the Solve stage alone, as
        # src/test/linear/main.cpp:102-108 times it
        rep = max(1, steps + warmup)
        _, it, res, sec = cpu.solve_reference(system, periodic=(False, False, False), tol=0.0,
                                              maxiter=MAXITER, block=32, threads=cores,
                                              repeat=rep, workdir=workdir)
        sample += "reference SolverConjugate, native backend, 32^3 blocks, OpenMP %d threads, best of %d" % (cores, rep)
        threads = cores
    else:
        t0 = time.perf_counter()
        _, it, res, _ = cpu.solve(system, periodic=(False, False, False), tol=0.0, maxiter=MAXITER)
        sec = time.perf_counter() - t0
        sample += "C restatement (oracle/cg_oracle.c), 1 thread"
        threads = 1
    value = n ** 3 * it / sec
    return {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
            "seconds_per_solve": sec, "iterations": it, "residual": res, "cells": n ** 3}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    base = run_reference_cpu(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["seconds_per_solve"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "reference CPU SolverConjugate on a bounded sample: " + base["sample"]},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------
def main_group(args):
    """The same step driven by ONE process over N GPUs through the in-process slab group
    (aphcg_group_*: what the aphros adapter uses for `cuda_devices N`).  Comparison record for
    the process-per-GPU launch the driver uses; prints its own JSON line ("mode": "group")."""
    from aphros_b200 import Conf, Mesh, SolverConjugateCudaGroup, capi, systems

    n = args.gpus
    if capi.device_count() < n:
        raise SystemExit("--group --gpus %d needs %d CUDA devices" % (n, n))
    shape = (args.size,) * 3 if args.strong else global_shape(n, args.size)
    nsph, seed = SEEDS[1] if args.strong else SEEDS.get(n, (512 * n, 20240610 + n))
    nsph = max(1, int(nsph * (args.size / 512.0) ** 3))
    cells = int(np.prod(shape))
    solver = SolverConjugateCudaGroup(Conf(tol=0.0, miniter=0, maxiter=MAXITER), {},
                                      Mesh(shape=shape, periodic=(False, False, False)), range(n))
    solver.AssembleSpheres(systems.random_spheres(nsph, seed))

    def step():
        solver.UploadGuess(None)
        return solver.Run()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = solver.LaunchCount()
    solver.TimerStart()
    iters, loop_ms = 0, 0.0
    for _ in range(args.steps):
        info = step()
        iters += info.iter
        loop_ms += info.loop_ms
    dev_ms = solver.TimerStop()
    clocks = sampler.stop()
    if not (np.isfinite(info.residual) and np.isfinite(info.residual0) and info.residual0 > 0):
        raise SystemExit("bench: non-finite residual (%r, initial %r)" % (info.residual, info.residual0))
    line = {
        "mode": "group", "residual": {"initial": info.residual0, "after_step": info.residual}, "metric": METRIC, "value": cells * iters / (dev_ms * 1e-3), "unit": UNIT,
        "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%dx%dx%d (nz,ny,nx), %d spheres 1000:1, Neumann walls, one process "
                               "driving %d z-slabs (aphcg_group_*), 101 iterations per step"
                               % (shape + (nsph, n)),
                   "cells": cells, "kernels": solver.Describe()},
        "loop_ms_per_step": loop_ms / args.steps,
        "gpu_launches": int(solver.LaunchCount() - launches0), "clocks": clocks,
    }
    if not args.no_e2e and cells * 88 < 48e9:
        rows = capi.PinnedArray(shape + (8,))
        x0 = capi.PinnedArray(shape)
        xs = capi.PinnedArray(shape)
        for q, (z0, nzl) in enumerate(solver.Slabs()):
            capi.check(capi.lib().aphcg_download_system(solver._member(q),
                                                        capi.ptr(rows.array[z0:z0 + nzl]), None))
        x0.array[...] = 0.0
        solver.Solve(rows.array, x0.array, xs.array)
        t0 = time.perf_counter()
        k = max(1, min(args.steps, 3))
        it2 = sum(solver.Solve(rows.array, x0.array, xs.array).iter for _ in range(k))
        sec = time.perf_counter() - t0
        line["e2e"] = {"value": cells * it2 / sec, "unit": UNIT, "h2d_bytes_per_step": cells * 72,
                       "d2h_bytes_per_step": cells * 8, "steps": k, "ms_per_step": sec * 1e3 / k,
                       "api": "aphcg_group_solve (C ABI), rank-wide pinned host arrays"}
        rows.free(), x0.free(), xs.free()
    print(json.dumps(line))
    solver.close()
    return 0


def main_converge(args, solver, shape, nsph, world, rank, dist):
    """BASELINE config 4 ("CG to 1e-8 relative residual"): ONE solve of the resident system
    from a zero guess to residual < reltol * initial residual.  The reference's Conf only has
    an absolute tolerance, so a first one-iteration run fetches the initial residual norm
    (aphcg_info.residual0).  Not the bench line: prints its own JSON record."""
    import torch
    from aphros_b200 import Conf
    cells = int(np.prod(shape))
    solver.SetConf(Conf(tol=0.0, miniter=0, maxiter=0))
    solver.UploadGuess(None)
    res0 = solver.Run().residual0
    if not (np.isfinite(res0) and res0 > 0):
        raise SystemExit("bench --converge: initial residual is %r" % res0)
    solver.SetConf(Conf(tol=args.converge * res0, miniter=0, maxiter=args.maxiter))
    solver.UploadGuess(None)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    solver.TimerStart()
    info = solver.Run()
    dev_ms = solver.TimerStop()
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            "mode": "converge", "metric": METRIC, "value": cells * info.iter / (dev_ms * 1e-3),
            "unit": UNIT, "n_gpus": world, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%dx%dx%d (nz,ny,nx), %d spheres, density jump %g:1, Neumann walls, "
                                   "zero guess, CG to %g x initial residual" % (shape + (nsph, args.contrast, args.converge)),
                       "cells": cells, "parallelism": "z-slab x%d" % world,
                       "kernels": solver.Describe()},
            "iterations": info.iter, "residual": info.residual, "residual0": res0,
            "relative_residual": info.residual / res0, "converged": bool(info.residual < args.converge * res0),
            "solve_ms": dev_ms, "ms_per_iteration": dev_ms / max(info.iter, 1)}))
    solver.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main_ours(args):
    import torch
    import torch.distributed as dist

    from aphros_b200 import Conf, SolverConjugateCuda, capi, distr, systems

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)"
                         % (args.gpus, world))
    if capi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)

    per_gpu = args.size
    if args.shape:   # tuning aid: explicit global (nz, ny, nx)
        shape = tuple(args.shape)
        nsph, seed = SEEDS[1]
    elif args.strong:  # BASELINE config 3: one size^3 domain cut into N z-slabs
        shape = (per_gpu, per_gpu, per_gpu)
        nsph, seed = SEEDS[1]
    else:
        shape = global_shape(world, per_gpu)
        nsph, seed = SEEDS.get(world, (512 * world, 20240610 + world))
    nsph = max(1, int(nsph * (per_gpu / 512.0) ** 3))
    if args.spheres is not None:
        nsph = args.spheres
    periodic = (False, False, False)
    mesh = distr.local_mesh(shape, periodic, rank, world, device=local_rank)
    conf = Conf(tol=0.0, miniter=0, maxiter=MAXITER)
    solver = SolverConjugateCuda(conf, {"jacobi_precond": args.precond}, mesh)
    if world > 1:
        distr.connect(solver)
    spheres = systems.random_spheres(nsph, seed)
    # system resident in HBM before the timed region
    solver.AssembleSpheres(spheres, rho_in=1.0 / args.contrast)
    cells_local = int(np.prod(mesh.local_shape))
    cells = int(np.prod(shape))
    if args.converge:
        return main_converge(args, solver, shape, nsph, world, rank, dist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        solver.UploadGuess(None)
        return solver.Run()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = solver.LaunchCount()
    barrier()
    solver.TimerStart()
    t0 = time.perf_counter()
    iters = 0
    loop_ms = 0.0
    for _ in range(args.steps):
        info = step()
        iters += info.iter
        loop_ms += info.loop_ms
    dev_ms = solver.TimerStop()
    if args.steps < 1:
        raise SystemExit("bench: --steps must be at least 1")
    if not (np.isfinite(info.residual) and np.isfinite(info.residual0) and info.residual0 > 0):
        raise SystemExit("bench: non-finite residual (%r, initial %r): the step did not compute"
                         % (info.residual, info.residual0))
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    launches = solver.LaunchCount() - launches0
    times = torch.tensor([dev_ms, wall_ms, loop_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, loop_ms = [float(v) for v in times.tolist()]
    value = cells * iters / (dev_ms * 1e-3)

    # ---- per-kernel timing of the same workload (CUDA events around each launch) ----
    roofline = None
    if world == 1:
        solver.UploadGuess(None)
        ms_dir, ms_upd = solver.ProfileKernels(30)
        peak, peak_src = peak_hbm_gbs()
        ach = B_ALG_DIR_SPMV * cells_local / (ms_dir * 1e-3) / 1e9
        roofline = {
            "bound": "hbm", "kernel": "k_dir_spmv (%s)" % solver.Describe().split(" ")[0], "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak,
            "traffic": measured_traffic(cells_local, solver.Describe().split(" ")[0]),
            "algorithmic_bytes_per_launch": B_ALG_DIR_SPMV * cells_local,
            "ms_per_launch": ms_dir, "peak_source": peak_src,
            "update_kernel": {"ms_per_launch": ms_upd,
                              "achieved": B_ALG_UPDATE * cells_local / (ms_upd * 1e-3) / 1e9},
            # what the HBM actually moved (ncu DRAM bytes of the committed capture / live time):
            # below the algorithmic figure because symmetric storage reads 4 of the 7 coefficients
            "dram": (None if not measured_traffic(cells_local, solver.Describe().split(" ")[0]) else {
                "bytes_per_launch": measured_traffic(cells_local, solver.Describe().split(" ")[0]),
                "achieved": measured_traffic(cells_local, solver.Describe().split(" ")[0]) / (ms_dir * 1e-3) / 1e9,
                "frac": measured_traffic(cells_local, solver.Describe().split(" ")[0]) / (ms_dir * 1e-3) / 1e9 / peak}),
            "iteration": {"algorithmic_bytes_per_cell": B_ALG,
                          "achieved": B_ALG * cells * iters / (loop_ms * 1e-3) / 1e9,
                          "frac": B_ALG * cells * iters / (loop_ms * 1e-3) / 1e9 / peak},
        }
    elif rank == 0:
        # several GPUs: the kernels cannot be timed one by one (the peers would wait), so the
        # roofline entry is the whole iteration, per GPU, against the same measured peak
        peak, peak_src = peak_hbm_gbs()
        ach = B_ALG * cells * iters / (loop_ms * 1e-3) / 1e9 / world
        roofline = {"bound": "hbm", "kernel": "whole CG iteration, per GPU (k_dir_spmv + k_update + hand-offs)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "algorithmic_bytes_per_cell": B_ALG, "peak_source": peak_src}

    # ---- end to end: host buffers in, host buffer out, through aphcg_solve -----------
    e2e = None
    if not args.no_e2e:
        rows = capi.PinnedArray(mesh.local_shape + (8,))
        x0 = capi.PinnedArray(mesh.local_shape)
        xs = capi.PinnedArray(mesh.local_shape)
        capi.check(capi.lib().aphcg_download_system(solver._h, capi.ptr(rows.array), None))
        x0.array[...] = 0.0
        e2e_steps = max(1, min(args.steps, 3))
        solver.Solve(rows.array, x0.array, xs.array)  # warm-up
        barrier()
        t0 = time.perf_counter()
        it2 = 0
        for _ in range(e2e_steps):
            it2 += solver.Solve(rows.array, x0.array, xs.array).iter
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        e2e = {"value": cells * it2 / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": cells * 72, "d2h_bytes_per_step": cells * 8,
               "steps": e2e_steps, "ms_per_step": e2e_s * 1e3 / e2e_steps,
               "api": "aphcg_solve (C ABI) with pinned host buffers: rows+guess H2D, solution D2H"}
        rows.free(), x0.free(), xs.free()

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            b = run_reference_cpu(1, 0)
            cpu_base = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the checker is optional for the bench line
            cpu_base = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable",
                        "sample": "failed: %s" % e}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": "%dx%dx%d (nz,ny,nx) synthetic variable-density Poisson (S3/S4: %d spheres, "
                            "1000:1 density jump, Neumann walls), %d^3 cells per GPU, z-slabs; one step = "
                            "one solve with tol=0 maxiter=%d (%d CG iterations, reference run_bench setting); "
                            "device assembly, zero guess" % (shape + (nsph, per_gpu, MAXITER, MAXITER + 1)),
                "cells": cells, "iterations_per_step": iters // max(args.steps, 1),
                "l2": "inputs (%.1f GB per GPU) far larger than the 126 MB L2; no flush needed"
                      % (cells_local * 8 * 13 / 1e9),
                "parallelism": "z-slab x%d" % world,
                "kernels": solver.Describe(),
            },
            "wall_ms_per_step": wall_ms / args.steps,
            "loop_ms_per_step": loop_ms / args.steps,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "residual": {"initial": info.residual0, "after_step": info.residual},
        }
        if roofline:
            line["roofline"] = roofline
        if e2e:
            line["e2e"] = e2e
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line))
    solver.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512, help="cells per GPU per direction")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--shape", type=int, nargs=3, default=None, metavar=("NZ", "NY", "NX"),
                    help="explicit global shape (kernel tuning)")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: one size^3 domain over all GPUs (default: size^3 per GPU)")
    ap.add_argument("--group", action="store_true",
                    help="one process drives all --gpus devices (in-process slab group) instead of "
                         "one process per GPU")
    ap.add_argument("--converge", type=float, default=0.0, metavar="RELTOL",
                    help="instead of the bench step: one solve to RELTOL x initial residual")
    ap.add_argument("--maxiter", type=int, default=100000, help="iteration limit of --converge")
    ap.add_argument("--precond", action="store_true",
                    help="opt-in Jacobi-preconditioned recurrence (not the reference's)")
    ap.add_argument("--spheres", type=int, default=None, help="number of spheres (0: constant density)")
    ap.add_argument("--contrast", type=float, default=1000.0, help="density jump outside:inside")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    if args.group:
        return main_group(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
