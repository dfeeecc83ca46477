#!/bin/bash
# round 2, call G (1 GPU): access-pattern microbenchmark, in-app timing (ap.mfer 202 at 64^3 and
# 128^3, conjugate vs conjugate_cuda), ncu launch list + full captures of the two loop kernels.
mkdir -p gpurun_out
nproc > gpurun_out/r2g_host.txt; free -g >> gpurun_out/r2g_host.txt; lscpu | grep -E "Model name|Socket|NUMA" >> gpurun_out/r2g_host.txt
timeout 300 scripts/stream_bench 2>&1 | tee gpurun_out/r2g_stream_bench.txt
timeout 900 python scripts/inapp_timing.py --sizes 64 128 --steps 3 --out gpurun_out/r2g_inapp_timing.json 2>&1 | tee gpurun_out/r2g_inapp_timing.md
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv \
  --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dir_spmv -s 4 -c 2 \
  -f -o gpurun_out/prof_dir_spmv $B > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_update -s 4 -c 2 \
  -f -o gpurun_out/prof_update $B >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -12
