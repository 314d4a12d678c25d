#!/bin/bash
# A/B of the 8-wide quantised BVH against the binary BVH on the two BVH-heavy configs, both pipelines.
mkdir -p gpurun_out
for cfg in c3 c4; do for pipe in regen wavefront; do for wide in 0 1; do
  echo "== $cfg $pipe wide=$wide" >> gpurun_out/wide_ab.log
  timeout 300 python bench.py --config $cfg --pipeline $pipe --bvh-wide $wide --no-cpu --steps 3 --warmup 3 --spp $([ $cfg = c3 ] && echo 16 || echo 64) 2>&1 | tail -1 | python -c "
import sys, json
l=sys.stdin.read().strip()
try:
    j=json.loads(l); print(j['value'], j['unit'], j['ms_per_step'], 'ms/step')
except Exception as e: print('ERR', l[-400:])
" >> gpurun_out/wide_ab.log
done; done; done
cat gpurun_out/wide_ab.log
