#!/bin/bash
# round 2, call Q (1 GPU): size sweep 128^3..1024^3 and the in-app timing with the final kernels.
mkdir -p gpurun_out
sed -i 's/--no-e2e --no-cpu >>/--no-e2e --no-cpu --no-parity >>/' scripts/gpu_sweep_sizes.sh
scripts/gpu_sweep_sizes.sh 64 96 128 192 256 384 512 768 1024 | tee gpurun_out/r2q_sweep_sizes.md
timeout 900 python scripts/inapp_timing.py --sizes 64 128 --steps 3 --out gpurun_out/r2q_inapp_timing.json 2>&1 | tee gpurun_out/r2q_inapp_timing.md
