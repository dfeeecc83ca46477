#!/bin/bash
# Sweep kernel configurations on the bench workload; one summary line each.
mkdir -p gpurun_out
out=gpurun_out/sweep.txt
: > $out
run() {
  echo "== $*" >> $out
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu 2>>gpurun_out/sweep.err | python -c "
import json,sys
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    r=j.get('roofline',{})
    print('value=%.4e ms/iter=%.4f dir_ms=%.4f upd_ms=%.4f it_frac=%.3f' % (j['value'], j['loop_ms_per_step']/101, r.get('ms_per_launch',0), r.get('update_kernel',{}).get('ms_per_launch',0), r.get('iteration',{}).get('frac',0)))
" >> $out
}
for cfg in "$@"; do run ${cfg//,/ }; done
cat $out
