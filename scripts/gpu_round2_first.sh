#!/bin/bash
# First GPU call of the next round (1 GPU, ~3 min): everything that was written after round 1's
# GPU budget ended.
#   1. the GPU tests that have never run: 2-D drop-in, example 201 in-app, config-1 captured
#      system, capture/replay, wide meshes (new tolerances);
#   2. the opt-in kDefer kernel variant: bitwise test, then even/odd kernel times with it off/on
#      at 512^3 and the small-mesh latencies.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf \
  -k "two_dimensional or taylor_couette or config1 or capture_and_replay or wide_meshes or device_resident" \
  > gpurun_out/r2_new_tests.log 2>&1
tail -15 gpurun_out/r2_new_tests.log
APHCG_TEST_DEFER=1 timeout 300 python -m pytest tests/test_gpu_z_late.py -q -rf -k deferred_consumption \
  > gpurun_out/r2_defer_test.log 2>&1
tail -5 gpurun_out/r2_defer_test.log
scripts/gpu_sweep_env.sh APHCG_DEFER=0 APHCG_DEFER=1 APHCG_DEFER=1,APHCG_PREFETCH=1 APHCG_DEFER=1,APHCG_PREFETCH=3
for d in 0 1; do
  echo "== small meshes, APHCG_DEFER=$d"
  APHCG_DEFER=$d timeout 300 python scripts/small_sweep.py X=1
done 2>&1 | tee gpurun_out/r2_small.txt
