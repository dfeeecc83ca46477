#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/sweep_env.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_round2.py -q -x -k "solution_parity or ragged or wide or bitwise or batched or fixed_iterations" 2>&1 | tail -3
scripts/gpu_sweep_env.sh X=1 X=2
timeout 300 ncu --metrics smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_dir_spmv -s 4 -c 2 --csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity 2>/dev/null | grep -E "k_dir_spmv" | cut -d, -f5,12- | head -4
