#!/bin/bash
# round 2, call A (1 GPU): the opt-in kDefer kernel variant -- bitwise test, then even/odd kernel
# times with it off/on at 512^3, and the small-mesh latencies.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt
APHCG_TEST_DEFER=1 timeout 300 python -m pytest tests/test_gpu_z_late.py -q -rf -k deferred_consumption \
  > gpurun_out/r2a_defer_test.log 2>&1
tail -5 gpurun_out/r2a_defer_test.log
scripts/gpu_sweep_env.sh APHCG_DEFER=0 APHCG_DEFER=1 APHCG_DEFER=1,APHCG_PREFETCH=1 APHCG_DEFER=1,APHCG_PREFETCH=3 APHCG_DEFER=1,APHCG_TILE=64
for d in 0 1; do
  echo "== small meshes, APHCG_DEFER=$d"
  APHCG_DEFER=$d timeout 300 python scripts/small_sweep.py X=1
done 2>&1 | tee gpurun_out/r2a_small.txt
