#!/bin/bash
# round 2, call C (2 GPUs): the in-kernel mailbox waits -- process-per-GPU parity in both wait
# modes, the slab group, bench.py at N=2 (parity key, strong key), config-4 mode at small scale.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -q -rP -k "not 4-mail and not 8-mail" \
  > gpurun_out/r2c_pytest_2gpu.log 2>&1
grep -E "passed|failed|OK|FAIL" gpurun_out/r2c_pytest_2gpu.log | tail -40
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
tail -c 2500 gpurun_out/r2c_bench_n2.json; tail -5 gpurun_out/r2c_bench_n2.err
APHCG_WAIT=finish timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2c_bench_n2_finish.json 2> gpurun_out/r2c_bench_n2_finish.err
tail -c 1200 gpurun_out/r2c_bench_n2_finish.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 2 --size 256 --converge 1e-8 --contrast 10 > gpurun_out/r2c_converge_n2.json 2> gpurun_out/r2c_converge_n2.err
cat gpurun_out/r2c_converge_n2.json; tail -3 gpurun_out/r2c_converge_n2.err
