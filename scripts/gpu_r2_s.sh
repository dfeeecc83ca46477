#!/bin/bash
# round 2, call S (1 GPU): round-2 tests after the last edits, the bench line with e2e_projection.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_round2.py -q -rP > gpurun_out/r2s_round2_tests.log 2>&1
grep -E "passed|failed|^E  |Error" gpurun_out/r2s_round2_tests.log | tail -12
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2s_bench_n1.json 2> gpurun_out/r2s_bench_n1.err
grep '^{' gpurun_out/r2s_bench_n1.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value %.4e e2e %.4e e2e_projection %s frac %.3f frac_dram %s parity %s' % (d['value'], d['e2e']['value'], json.dumps(d.get('e2e_projection'))[:400], d['roofline']['frac'], d['roofline']['frac_dram'], d['parity']['ok']))
"; tail -6 gpurun_out/r2s_bench_n1.err
