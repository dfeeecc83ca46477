#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
  bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err
echo "rc=$?"
grep '^{' gpurun_out/r2w_bench_n2.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value %.4e e2e %.4e e2e_projection %.4e (%.0f ms) parity %s strong eff %.3f' % (d['value'], d['e2e']['value'], d['e2e_projection']['value'], d['e2e_projection']['ms_per_step'], d['parity']['ok'], d['strong']['efficiency_vs_n1']))
"; tail -4 gpurun_out/r2w_bench_n2.err
