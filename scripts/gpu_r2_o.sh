#!/bin/bash
# round 2, call O (2 GPUs): the stream kernel in the multi-rank paths: parity tests (process per
# GPU, both wait modes, NCCL; slab group), bench N=2 with parity + strong keys.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -q -rP -k "not 4-mail and not 8-mail" \
  > gpurun_out/r2o_pytest_2gpu.log 2>&1
grep -E "passed|failed|rank 0 .*(OK|FAIL)" gpurun_out/r2o_pytest_2gpu.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err
grep '^{' gpurun_out/r2o_bench_n2.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value %.4e e2e %.4e loop/iter %.4f parity %s strong %s numa %s' % (d['value'], d['e2e']['value'], d['loop_ms_per_step']/101, d['parity']['ok'], json.dumps({k: d['strong'][k] for k in ('ms_per_iteration','n1_ms_per_iteration','efficiency_vs_n1')}), d['config']['numa_binding']))
"; tail -3 gpurun_out/r2o_bench_n2.err
