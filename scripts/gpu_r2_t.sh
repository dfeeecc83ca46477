#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_round2.py -q -k "bitwise or connections" > gpurun_out/r2t_tests.log 2>&1; tail -3 gpurun_out/r2t_tests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_small.py > gpurun_out/r2t_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/r2t_memcheck.log
