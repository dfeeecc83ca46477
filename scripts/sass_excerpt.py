"""profiles/<tag>_sass_tma.txt: what the shipped libaphcg.so contains, from `cuobjdump -sass`
(no GPU needed): per kernel the counts of the TMA / mbarrier / L2-prefetch / 128-bit memory
instructions, then those lines of the benchmark's direction kernel in program order.

    python scripts/sass_excerpt.py r02
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "aphros_b200", "libaphcg.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
PAT = re.compile(r"\b(UTMALDG\.\dD|UTMAPF\S*|SYNCS\.[A-Z0-9.]+|UBLKPF\.L2|LDG\.E\.[A-Z.]*128\S*|"
                 r"STG\.E\.[A-Z.]*128\S*|LDS\.128|STS\.128|MEMBAR\.[A-Z.]+|CCTL\.IVALL|DFMA|HMMA\S*|UTCMMA\S*)")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
elfs = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
counts, lines, fn = collections.OrderedDict(), collections.defaultdict(list), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(anonymous namespace\)::", "", fn).split("(")[0].replace("void acg::", "")
        counts[fn] = collections.Counter()
        continue
    m = PAT.search(line)
    if fn and m:
        counts[fn][m.group(1)] += 1
        lines[fn].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).strip())
out = os.path.join(ROOT, "profiles", tag + "_sass_tma.txt")
with open(out, "w") as f:
    f.write("# cuobjdump -sass aphros_b200/libaphcg.so (python scripts/sass_excerpt.py %s)\n" % tag)
    f.write("# embedded cubins (all sm_100a):\n" + "".join("#   %s\n" % l for l in elfs.splitlines()))
    f.write("\n# per kernel: TMA loads (UTMALDG), mbarrier ops (SYNCS.*), L2 prefetch (UBLKPF.L2), "
            "128-bit memory instructions, fences, FP64 FMAs; no tensor-core instructions anywhere\n")
    for fn, c in counts.items():
        if c:
            f.write("%-46s %s\n" % (fn, "  ".join("%s x%d" % kv for kv in sorted(c.items()))))
    key = next(k for k in counts if k.startswith("k_dir_spmv_stream<64, true>"))
    f.write("\n# %s -- the benchmark's direction kernel, in program order\n" % key)
    f.write("\n".join(lines[key]) + "\n")
print("wrote", out)
