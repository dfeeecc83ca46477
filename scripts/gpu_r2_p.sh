#!/bin/bash
# round 2, call P (8 GPUs): the final code at N=8 -- bench line (parity, strong, NUMA binding),
# 8-rank parity tests, host topology.
mkdir -p gpurun_out
{ nvidia-smi topo -m; numactl -H 2>/dev/null || lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)"; free -g | head -2; } > gpurun_out/r2p_topology.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 \
  > gpurun_out/r2p_bench_n8.json 2> gpurun_out/r2p_bench_n8.err
echo "bench rc=$?"
grep '^{' gpurun_out/r2p_bench_n8.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value %.4e e2e %.4e loop/iter %.4f parity %s strong %s numa %s clocks %s' % (d['value'], d['e2e']['value'], d['loop_ms_per_step']/101, json.dumps({k: (d['parity'][k].get('iter'), d['parity'][k].get('rel_max_abs'), d['parity'][k].get('ok')) for k in ('tlinear','walls')}), json.dumps({k: d['strong'][k] for k in ('ms_per_iteration','n1_ms_per_iteration','efficiency_vs_n1')}), d['config']['numa_binding'], d['clocks']))
"; tail -3 gpurun_out/r2p_bench_n8.err
timeout 600 python -m pytest tests/test_gpu_multi.py -q -rP -k "8-mail" > gpurun_out/r2p_pytest_8ranks.log 2>&1
grep -E "passed|failed|rank 0 .*(OK|FAIL)" gpurun_out/r2p_pytest_8ranks.log | tail -8
