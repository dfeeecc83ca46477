#!/bin/bash
# round 2, last call (1 GPU): whole GPU suite and the default bench line with the final code.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rf > gpurun_out/r2z_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2z_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/r2z_bench_default.json 2> gpurun_out/r2z_bench_default.err
grep '^{' gpurun_out/r2z_bench_default.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value %.4e e2e %.4e proj %.4e frac %.3f frac_dram %.3f iter_frac %.3f parity %s launches %d clocks %s' % (d['value'], d['e2e']['value'], d['e2e_projection']['value'], d['roofline']['frac'], d['roofline']['frac_dram'], d['roofline']['iteration']['frac'], d['parity']['ok'], d['gpu_launches'], d['clocks']))
"; tail -2 gpurun_out/r2z_bench_default.err
