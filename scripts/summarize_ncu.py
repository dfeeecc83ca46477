#!/usr/bin/env python
"""Turns gpurun_out/ ncu artefacts into small tracked summaries under profiles/.

  python scripts/summarize_ncu.py <tag>     (e.g. r01b)

  gpurun_out/launches.csv         -> profiles/<tag>_launches.md   (per-kernel totals/shares)
  gpurun_out/prof_*.ncu-rep       -> profiles/<tag>_<name>_raw.csv (selected raw metrics)
"""
import collections
import csv
import glob
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]

lp = os.path.join(GO, "launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    with open(os.path.join(OUT, tag + "_launches.md"), "w") as f:
        f.write("# ncu launch list (%s): `ncu --metrics gpu__time_duration.sum --clock-control none "
                "python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu`\n\n" % tag)
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write("| `%s` | %d | %.3f | %.1f | %.1f%% |\n" % (k, cnt[k], v / 1e6, v / cnt[k] / 1e3, 100 * v / T))
    print("wrote", tag + "_launches.md")

for rep in sorted(glob.glob(os.path.join(GO, "prof_*.ncu-rep"))):
    name = os.path.basename(rep)[5:-8]
    p = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(p.stdout.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    with open(os.path.join(OUT, "%s_%s_raw.csv" % (tag, name)), "w") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "metric", "value", "unit"])
        for r in rows[2:]:
            for m in WANT:
                if m in hdr:
                    i = hdr.index(m)
                    w.writerow([r[hdr.index("ID")], r[hdr.index("Kernel Name")][:48], m, r[i], units[i]])
    print("wrote", "%s_%s_raw.csv" % (tag, name))

# ---- <tag>_traffic.json: DRAM bytes per launch of the two loop kernels (bench.py's
# roofline.traffic); usage: summarize_ncu.py <tag> <kernel tag of aphcg_describe> [cells]
if len(sys.argv) > 2:
    import json
    ktag = sys.argv[2]
    cells = int(sys.argv[3]) if len(sys.argv) > 3 else 512 ** 3

    def dram_bytes(name):
        path = os.path.join(OUT, "%s_%s_raw.csv" % (tag, name))
        per = collections.defaultdict(float)
        for r in csv.DictReader(open(path)):
            if r["metric"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[r["unit"]]
                per[int(r["launch"])] += float(r["value"]) * scale
        return [per[k] for k in sorted(per)]

    d, u = dram_bytes("dir_spmv"), dram_bytes("update")
    rec = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full "
                       "capture summarised in %s_*_raw.csv (two consecutive launches of each kernel: "
                       "with batched x updates the direction kernel alternates between an 'even' launch "
                       "that applies two deferred x updates and an 'odd' one that applies none); "
                       "bench.py reports the mean as roofline.traffic" % tag,
           "cells": cells,
           "kernels": {ktag: {"k_dir_spmv_bytes_per_launch": sum(d) / len(d),
                              "k_dir_spmv_launches": d,
                              "k_update_bytes_per_launch": sum(u) / len(u)}}}
    with open(os.path.join(OUT, tag + "_traffic.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print("wrote", tag + "_traffic.json", rec["kernels"][ktag]["k_dir_spmv_bytes_per_launch"] / cells, "B/cell")
