"""Warp-stall summary of a full ncu capture (source page, SASS level): stall-reason totals,
thread instructions per cell, and the instructions where the samples sit.

    python scripts/stall_summary.py gpurun_out/prof_dir_spmv.ncu-rep 134217728 > profiles/r02_dir_spmv_stalls.md
"""
import collections
import csv
import subprocess
import sys

rep, cells = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
kernels, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and row and row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] and len(row) == len(cur["hdr"]):
        cur["rows"].append(row)
for n, k in enumerate(kernels):
    h = k["hdr"]
    col = {name: i for i, name in enumerate(h)}
    stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    tot = collections.Counter()
    thread_inst = warp_inst = samples = 0
    for r in k["rows"]:
        for c in stall_cols:
            tot[c] += int(r[col[c]] or 0)
        thread_inst += int(r[col["Thread Instructions Executed"]] or 0)
        warp_inst += int(r[col["Instructions Executed"]] or 0)
        samples += int(r[col["# Samples"]] or 0)
    print("## launch %d: %s" % (n, k["name"].split("(")[0][:90]))
    print()
    print("%d SASS instructions, %.3e warp instructions, %.3e thread instructions = **%.1f per cell**, %d samples"
          % (len(k["rows"]), warp_inst, thread_inst, thread_inst / cells, samples))
    print()
    print("stall reasons: " + ", ".join("%s %.1f%%" % (c[6:], 100.0 * v / max(samples, 1))
                                        for c, v in tot.most_common(9)))
    print()
    print("| share of samples | SASS | top reasons |")
    print("|---|---|---|")
    rows = sorted(k["rows"], key=lambda r: -int(r[col["# Samples"]] or 0))[:14]
    for r in rows:
        s = int(r[col["# Samples"]] or 0)
        rs = sorted(((int(r[col[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print("| %.1f%% | `%s` | %s |" % (100.0 * s / max(samples, 1), r[col["Source"]].strip()[:70],
                                          ", ".join("%s %d%%" % (c, 100 * v / max(s, 1)) for v, c in rs if v)))
    print()
