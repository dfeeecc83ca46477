#!/bin/bash
# One 8-GPU call: weak-scaling bench line (process per GPU), the same step through the
# in-process slab group, a config-4 style solve to 1e-8 relative residual at 1024^3, and the
# 8-GPU group tests.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
timeout 300 $TR 29501 bench.py --gpus 8 --steps 2 --warmup 3 --no-e2e > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
cat gpurun_out/bench_n8.json
timeout 300 python bench.py --group --gpus 8 --steps 2 --warmup 3 --no-e2e > gpurun_out/bench_group8.json 2> gpurun_out/bench_group8.err
cat gpurun_out/bench_group8.json
timeout 400 $TR 29502 bench.py --gpus 8 --converge 1e-8 --contrast 10 --maxiter 40000 > gpurun_out/converge_n8.json 2> gpurun_out/converge_n8.err
cat gpurun_out/converge_n8.json
timeout 300 python -m pytest tests/test_gpu_group.py -q -x -k 8gpus 2>&1 | tail -3
tail -3 gpurun_out/bench_n8.err gpurun_out/bench_group8.err gpurun_out/converge_n8.err
