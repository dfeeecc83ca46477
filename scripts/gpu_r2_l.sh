#!/bin/bash
# round 2, call L (1 GPU): the all-operands-by-TMA variant of the direction kernel.
mkdir -p gpurun_out
APHCG_STREAM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "solution_parity or batched or nonsymmetric or ragged or wide or fixed_iterations or deterministic or tlinear_32" > gpurun_out/r2l_stream_tests.log 2>&1
tail -15 gpurun_out/r2l_stream_tests.log
rm -f gpurun_out/sweep_env.txt
scripts/gpu_sweep_env.sh X=1 APHCG_STREAM=1 APHCG_STREAM=1,APHCG_PREFETCH=0 APHCG_STREAM=1,APHCG_PREFETCH=2 APHCG_STREAM=1,APHCG_ZC=16 APHCG_STREAM=1,APHCG_ZC=64
