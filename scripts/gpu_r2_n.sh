#!/bin/bash
# round 2, call N (1 GPU): the whole GPU suite with the stream kernel as default, the bench line,
# ncu launch list + full captures of the final kernels, smoke.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/r2n_pytest_gpu.log 2>&1
tail -8 gpurun_out/r2n_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err
cut -c1-400 gpurun_out/r2n_bench_n1.json; tail -3 gpurun_out/r2n_bench_n1.err
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/r2n_smoke.txt
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv \
  --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dir_spmv -s 4 -c 2 \
  -f -o gpurun_out/prof_dir_spmv $B > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_update -s 4 -c 2 \
  -f -o gpurun_out/prof_update $B >> gpurun_out/ncu_full.log 2>&1
APHCG_PERSISTENT=0 timeout 300 python scripts/small_sweep.py X=1 2>&1 | tee gpurun_out/r2n_small_stream.txt
