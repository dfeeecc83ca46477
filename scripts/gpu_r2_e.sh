#!/bin/bash
# round 2, call E (2 GPUs): who waits for the all-reduced scalars -- finish kernels, the consumer
# kernels' CTAs (LL mailboxes, speculative peek), the same without the reader fence (measurement
# only) -- on 64-plane slabs (what each of 8 GPUs owns in the 512^3 strong-scaling run) and on
# the 512^3-per-GPU weak-scaling step.
mkdir -p gpurun_out
run() {
  local tag=$1; shift
  echo "== $tag"
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-strong $SHAPE 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.4e  loop_ms/iter %.4f  parity %s  %s' % (d['value'], d['loop_ms_per_step']/101, d.get('parity',{}).get('ok'), d['config']['kernels'].split('allreduce=')[-1]))
"
}
{
for SHAPE in "--shape 128 512 512" ""; do
  echo "#### ${SHAPE:-weak 512^3 per GPU}"
  run finish APHCG_WAIT=finish
  run kernel APHCG_WAIT=kernel
  run kernel-nofence APHCG_WAIT=kernel APHCG_WAIT_FENCE=0
  run nccl APHCG_ALLREDUCE=nccl
done
} 2>&1 | tee gpurun_out/r2e_wait_modes.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q -rP -k "2-mail" > gpurun_out/r2e_pytest_2gpu.log 2>&1
grep -E "passed|failed|OK|FAIL" gpurun_out/r2e_pytest_2gpu.log | tail -12
