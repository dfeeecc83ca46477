#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/r2v_pytest_gpu.log 2>&1
tail -6 gpurun_out/r2v_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
