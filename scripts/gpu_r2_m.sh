#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/sweep_env.txt
scripts/gpu_sweep_env.sh APHCG_STREAM=1 APHCG_STREAM=1,APHCG_ZC=16 APHCG_STREAM=1,APHCG_ZC=64 APHCG_STREAM=1,APHCG_ZC=128 APHCG_STREAM=1,APHCG_PSTREAM=0 APHCG_STREAM=1,APHCG_UPD_CTAS=8 APHCG_STREAM=1
for sh in "64 512 512" "256 256 256" "128 128 128" "192 192 192"; do
for cfg in APHCG_STREAM=0 APHCG_STREAM=1; do
  echo "== shape $sh $cfg"
  env $cfg APHCG_VERBOSE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --shape $sh 2>&1 | grep -E "aphcg profile" | cut -c1-200
done; done | tee -a gpurun_out/sweep_env.txt
APHCG_STREAM=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2m_bench_stream.json 2>/dev/null; cut -c1-700 gpurun_out/r2m_bench_stream.json
