#!/bin/bash
# usage: gpu_sweep_env.sh cfg1 cfg2 ...   (cfg = comma-separated env assignments)
# prints the per-kernel profile line (even/odd direction-kernel launches) of the 512^3 bench system
mkdir -p gpurun_out
for cfg in "$@"; do
  echo "== $cfg"
  env ${cfg//,/ } APHCG_VERBOSE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | grep -E "aphcg profile" | cut -c1-200
done | tee -a gpurun_out/sweep_env.txt
