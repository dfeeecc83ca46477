"""In-app timing (GPU box): the reference's own application ap.mfer runs
examples/202_coalescence for a few time steps, once with `linsolver_symm conjugate` (the
reference's CPU CG, OpenMP over blocks on all host cores) and once with the CUDA module
preloaded and selected by name; the stage timers the application prints itself
(`verbose_time` / `verbose_stages`, src/distr/distr.ipp:581-630) are parsed into a table:
the projection's pressure solve (`project:01:solve`), the velocity solves
(`diffusion:*:solve`), and -- for conjugate_cuda -- what the adapter adds around the C ABI
call (its own stages gather / solve / scatter).

    python scripts/inapp_timing.py [--sizes 64 128] [--steps 3] [--out gpurun_out/inapp.json]
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
PLUGIN = os.path.join(ROOT, "aphros_b200", "plugin", "libaphcg_aphros.so")
STAGE = re.compile(r"^([| ]*)(\S*) \[(\d+\.\d+) s, ")


def run(size, solver, steps, threads, extra=""):
    d = tempfile.mkdtemp(prefix="inapp_%d_%s_" % (size, solver))
    shutil.rmtree(d)
    shutil.copytree(os.path.join(REF, "app202"), d)
    nb = size // 32
    with open(os.path.join(d, "mesh.conf"), "w") as f:
        f.write("".join("set int %s %d\n" % kv for kv in [
            ("px", 1), ("py", 1), ("pz", 1), ("bx", nb), ("by", nb), ("bz", nb),
            ("bsx", 32), ("bsy", 32), ("bsz", 32)]))
    with open(os.path.join(d, "add.conf"), "w") as f:
        f.write("set string linsolver_symm %s\n" % solver)
        f.write("set string linsolver_gen conjugate\nset string linsolver_vort conjugate\n")
        f.write("set int max_step %d\nset double tmax 100\nset int linreport 0\n" % steps)
        f.write("set int dumppoly 0\nset string dumplist\nset double dump_field_dt 1e10\n")
        f.write("set int verbose_time 1\nset int verbose_stages 1\n")
        f.write(extra)
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    if solver != "conjugate":
        env["LD_PRELOAD"] = PLUGIN
    p = subprocess.run([os.path.join(REF, "ap.mfer"), "a.conf"], cwd=d, env=env,
                       capture_output=True, text=True, timeout=3000)
    if p.returncode != 0:
        raise SystemExit("ap.mfer failed:\n" + p.stdout[-2000:] + p.stderr[-2000:])
    # accumulate the report's times by stage name, and by (parent, name) for the solver stages
    acc, path = {}, []
    for line in p.stderr.splitlines():
        m = STAGE.match(line)
        if not m:
            continue
        depth = m.group(1).count("|")
        name, sec = m.group(2), float(m.group(3))
        path[depth:] = [name]
        acc[name] = acc.get(name, 0.0) + sec
        if name.startswith("Solve:") and depth > 0:
            key = path[depth - 1].split(":")[0] + "/" + name
            acc[key] = acc.get(key, 0.0) + sec
    total = float(re.search(r"total = (\S+) s", p.stderr).group(1))
    shutil.rmtree(d, ignore_errors=True)
    return {"total_s": total,
            "pressure_solve_s": acc.get("project:01:solve", 0.0),
            "velocity_solves_s": sum(v for k, v in acc.items()
                                     if re.match(r"diffusion:\d+:solve$", k)),
            "project_local_s": acc.get("project:00:local", 0.0),
            "solve_stages_s": {k: v for k, v in sorted(acc.items()) if "/Solve:" in k}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[64, 128])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "inapp_timing.json"))
    args = ap.parse_args()
    threads = os.cpu_count() or 1
    rec = {"example": "examples/202_coalescence, %d time steps, 32^3 blocks, 1 rank, OpenMP %d threads"
                      % (args.steps, threads), "runs": []}
    print("| mesh | solver | whole run | pressure solves (project:01:solve) | velocity solves | "
          "adapter: gather | C ABI call | scatter |")
    print("|---|---|---|---|---|---|---|---|")
    for n in args.sizes:
        for solver in ("conjugate", "conjugate_cuda"):
            r = run(n, solver, args.steps, threads)
            r.update(mesh=n, solver=solver)
            rec["runs"].append(r)
            st = r["solve_stages_s"]
            g = sum(v for k, v in st.items() if k.endswith(":gather"))
            s = sum(v for k, v in st.items() if k.endswith(":solve"))
            c = sum(v for k, v in st.items() if k.endswith(":scatter"))
            print("| %d^3 | %s | %.3f s | %.3f s | %.3f s | %s | %s | %s |" % (
                n, solver, r["total_s"], r["pressure_solve_s"], r["velocity_solves_s"],
                *(("%.3f s" % v if solver != "conjugate" else "-") for v in (g, s, c))), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__":
    sys.exit(main())
