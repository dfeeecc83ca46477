#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_bench_reference.err
echo "rc=$?"; cut -c1-900 gpurun_out/r2x_bench_reference.json; tail -4 gpurun_out/r2x_bench_reference.err
