#!/bin/bash
# round 2, call F (8 GPUs): BASELINE config 4 (1024^3, CG to 1e-8 relative), the bench line at
# N=8 with its parity and strong-scaling keys, the multi-GPU parity tests at 2/4/8 ranks with
# their logs, then config 4 at 1000:1 with the opt-in preconditioner.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2f_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29531 bench.py --gpus 8 --converge 1e-8 --contrast 10 \
  > gpurun_out/r2f_config4_plain_10to1.json 2> gpurun_out/r2f_config4_plain.err
cat gpurun_out/r2f_config4_plain_10to1.json; tail -2 gpurun_out/r2f_config4_plain.err
timeout 400 $TR --master-port 29532 bench.py --gpus 8 --steps 5 --warmup 3 \
  > gpurun_out/r2f_bench_n8.json 2> gpurun_out/r2f_bench_n8.err
tail -c 3000 gpurun_out/r2f_bench_n8.json; tail -2 gpurun_out/r2f_bench_n8.err
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -q -rP \
  > gpurun_out/r2f_pytest_multigpu.log 2>&1
grep -E "passed|failed|rank 0 .*(OK|FAIL)" gpurun_out/r2f_pytest_multigpu.log | tail -40
timeout 600 $TR --master-port 29533 bench.py --gpus 8 --converge 1e-8 --contrast 1000 --precond \
  > gpurun_out/r2f_config4_precond_1000to1.json 2> gpurun_out/r2f_config4_precond.err
cat gpurun_out/r2f_config4_precond_1000to1.json; tail -2 gpurun_out/r2f_config4_precond.err
