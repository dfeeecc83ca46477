#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/sweep_env.txt
timeout 300 python -m pytest tests/test_gpu_zz_round2.py -q -k "bitwise" 2>&1 | tail -2
scripts/gpu_sweep_env.sh X=1 APHCG_UPD_CTAS=8 APHCG_UPD_CTAS=12 APHCG_UPD_CTAS=24 APHCG_UPD_UR=4 APHCG_UPD_UR=4,APHCG_UPD_CTAS=32
for cfg in X=1 APHCG_ZC=16 APHCG_ZC=22 APHCG_ZC=64; do
  echo "== shape 64 512 512 $cfg"
  env $cfg APHCG_VERBOSE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --shape 64 512 512 2>&1 | grep -E "aphcg profile" | cut -c1-200
done | tee -a gpurun_out/sweep_env.txt
