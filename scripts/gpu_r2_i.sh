#!/bin/bash
# round 2, call I (1 GPU): new tests (projection entry through the adapter, persistent loop),
# the reference arm at 512^3, the default bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_round2.py -q -rP > gpurun_out/r2i_round2_tests.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/r2i_round2_tests.log | tail -20
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r2i_bench_reference.json 2> gpurun_out/r2i_bench_reference.err
cat gpurun_out/r2i_bench_reference.json | cut -c1-1200; tail -5 gpurun_out/r2i_bench_reference.err
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err
cat gpurun_out/r2i_bench_n1.json; tail -5 gpurun_out/r2i_bench_n1.err
