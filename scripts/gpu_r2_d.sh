#!/bin/bash
# round 2, call D (1 GPU): planes-per-CTA of the direction kernel on the slab shapes of the
# strong-scaling runs (a 64- or 128-plane slab of 512^2 is what one of 8 / 4 GPUs owns), and
# on the mid-size meshes; the round-2 tests written so far.
mkdir -p gpurun_out
run() { # shape..., env
  local sh="$1 $2 $3"; shift 3
  for cfg in "$@"; do
    echo "== shape $sh  $cfg"
    env ${cfg//,/ } APHCG_VERBOSE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity \
      --shape $sh 2>&1 | grep -E "aphcg profile|\"value\"" | sed -E 's/.*"value": ([0-9.e+]+).*"ms_per_step": ([0-9.]+).*/value \1 ms_per_step \2/' | cut -c1-220
  done
}
{
run 64 512 512 X=1 APHCG_ZC=16 APHCG_ZC=64 APHCG_ZC=22 APHCG_ZC=11
run 128 512 512 X=1 APHCG_ZC=16 APHCG_ZC=64 APHCG_ZC=43
run 192 192 192 X=1 APHCG_ZC=16 APHCG_ZC=24 APHCG_ZC=48 APHCG_ZC=64
run 128 128 128 X=1 APHCG_ZC=16 APHCG_ZC=4 APHCG_TILE=64 APHCG_TILE=64,APHCG_ZC=4 APHCG_TILE=64,APHCG_ZC=16
run 256 256 256 X=1 APHCG_ZC=16 APHCG_ZC=64
} 2>&1 | tee gpurun_out/r2d_zc.txt
timeout 900 python -m pytest tests/test_gpu_zz_round2.py tests/test_gpu_fullsize.py -q -rP -k "round2 or config5" > gpurun_out/r2d_tests.log 2>&1
grep -E "passed|failed|config 5|1000:1|Error|assert" gpurun_out/r2d_tests.log | tail -20
