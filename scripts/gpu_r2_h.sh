#!/bin/bash
# round 2, call H (1 GPU): the persistent small-mesh loop -- its tests, the small-mesh sweep with
# it on and off, smoke(); then the whole GPU suite.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_round2.py -q -rP -k persistent > gpurun_out/r2h_persistent_tests.log 2>&1
grep -E "passed|failed|Error|assert " gpurun_out/r2h_persistent_tests.log | tail -20
for p in 0 1; do
  echo "== small meshes, APHCG_PERSISTENT=$p"
  APHCG_PERSISTENT=$p timeout 300 python scripts/small_sweep.py X=1
done 2>&1 | tee gpurun_out/r2h_small.txt
APHCG_PERSISTENT=100000000 timeout 300 python scripts/small_sweep.py X=1 2>&1 | tee -a gpurun_out/r2h_small.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/r2h_smoke.txt
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/r2h_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2h_pytest_gpu.log
