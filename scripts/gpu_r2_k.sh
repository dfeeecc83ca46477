#!/bin/bash
mkdir -p gpurun_out
scripts/gpu_sweep_env.sh APHCG_RPT=1 APHCG_TILE=64,APHCG_RPT=1 APHCG_TILE=64 APHCG_RPT=2
for cfg in APHCG_RPT=2 APHCG_RPT=1; do
  echo "== shape 64 512 512 $cfg"
  env $cfg APHCG_VERBOSE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --shape 64 512 512 2>&1 | grep -E "aphcg profile" | cut -c1-200
  echo "== shape 256 256 256 $cfg"
  env $cfg APHCG_VERBOSE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --shape 256 256 256 2>&1 | grep -E "aphcg profile" | cut -c1-200
done | tee -a gpurun_out/sweep_env.txt
