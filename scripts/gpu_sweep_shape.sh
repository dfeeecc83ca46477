#!/bin/bash
# usage: gpu_sweep_shape.sh "NZ NY NX" cfg1 cfg2 ...   (cfg = comma-separated env assignments)
mkdir -p gpurun_out
shape=$1; shift
out=gpurun_out/sweep_shape.txt
for cfg in "$@"; do
  echo "== shape $shape :: $cfg" >> $out
  env ${cfg//,/ } timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --shape $shape 2>>gpurun_out/sweep.err | python -c "
import json,sys
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    r=j.get('roofline',{})
    print('value=%.4e ms/iter=%.4f dir_ms=%.4f upd_ms=%.4f %s' % (j['value'], j['loop_ms_per_step']/101, r.get('ms_per_launch',0), r.get('update_kernel',{}).get('ms_per_launch',0), j['config']['kernels']))
" >> $out
done
cat $out
