#!/bin/bash
# round 2, call R (8 GPUs): BASELINE config 4 with the final kernels.
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
  bench.py --gpus 8 --converge 1e-8 --contrast 10 > gpurun_out/r2r_config4_plain_10to1.json 2> gpurun_out/r2r_config4.err
cat gpurun_out/r2r_config4_plain_10to1.json; tail -2 gpurun_out/r2r_config4.err
