#!/bin/bash
# Throughput sweep over mesh sizes on ONE GPU (north star: 128^3 .. 1024^3):
# one bench line per size into gpurun_out/sweep_sizes.jsonl, and a table on stdout.
# usage: scripts/gpu_sweep_sizes.sh [sizes...]
mkdir -p gpurun_out
out=gpurun_out/sweep_sizes.jsonl
: > $out
sizes=${@:-128 192 256 384 512 768 1024}
for n in $sizes; do
  timeout 600 python bench.py --size $n --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity >> $out 2>> gpurun_out/sweep_sizes.err
done
python - <<'PY'
import json
print("| size | cells | ms/iteration | cell-iter/s | 144 B/cell-iter GB/s | frac of copy peak | dir+SpMV ms | update ms | kernels |")
print("|---|---|---|---|---|---|---|---|---|")
for l in open("gpurun_out/sweep_sizes.jsonl"):
    try:
        j = json.loads(l)
    except Exception:
        continue
    r = j["roofline"]
    it = j["config"]["iterations_per_step"]
    print("| %s | %.3g | %.4f | %.3e | %.0f | %.2f | %.4f | %.4f | %s |" % (
        j["config"]["workload"].split(" ")[0], j["config"]["cells"], j["loop_ms_per_step"] / it,
        j["value"], r["iteration"]["achieved"], r["iteration"]["frac"], r["ms_per_launch"],
        r["update_kernel"]["ms_per_launch"], j["config"]["kernels"].split(" ctas")[0]))
PY
