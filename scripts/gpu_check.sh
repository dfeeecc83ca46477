#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list + full capture.
# usage: scripts/gpu_check.sh [tests|bench|ncu|all]
set -u
what=${1:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
if [ "$what" = tests ] || [ "$what" = all ]; then
  timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_gpu.log 2>&1
  grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_gpu.log | head -60
fi
if [ "$what" = bench ] || [ "$what" = all ]; then
  timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  APHCG_SPMV=plain timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_plain.json 2>> gpurun_out/bench.err
  cat gpurun_out/bench_plain.json
fi
if [ "$what" = ncu ] || [ "$what" = all ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu \
    > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dir_spmv -s 4 -c 2 \
    -f -o gpurun_out/prof_dir_spmv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu \
    > gpurun_out/ncu_full.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_update -s 4 -c 2 \
    -f -o gpurun_out/prof_update python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu \
    >> gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
