#!/bin/bash
# round 2, call J (1 GPU): the 512-thread (one row pair per thread) variant of the direction kernel.
mkdir -p gpurun_out
cat /sys/fs/cgroup/memory.max /sys/fs/cgroup/memory.high 2>/dev/null | tr '\n' ' ' > gpurun_out/r2j_cgroup.txt; echo >> gpurun_out/r2j_cgroup.txt
APHCG_RPT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "solution_parity or batched or nonsymmetric or ragged or wide or fixed_iterations or deterministic" > gpurun_out/r2j_rpt1_tests.log 2>&1
tail -3 gpurun_out/r2j_rpt1_tests.log
scripts/gpu_sweep_env.sh X=1 APHCG_RPT=1 APHCG_RPT=1,APHCG_PREFETCH=1 APHCG_RPT=1,APHCG_PREFETCH=3 APHCG_RPT=1,APHCG_ZC=16 APHCG_RPT=1,APHCG_PSTREAM=0
cat gpurun_out/r2j_cgroup.txt
