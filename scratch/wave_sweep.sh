#!/bin/bash
# paths per wavefront wave (RTB200_WAVE_MPATHS, default 8 M) on the two BVH-heavy configs
mkdir -p gpurun_out; L=gpurun_out/wave_sweep.log; rm -f $L
run() { echo "== WAVE=$RTB200_WAVE_MPATHS $*" >> $L; timeout 200 python bench.py "$@" --no-cpu --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
l=sys.stdin.read().strip()
try:
    j=json.loads(l); print(j['value'], j['unit'], j['ms_per_step'], 'ms/step')
except Exception as e: print('ERR', l[-300:])
" >> $L; }
for w in 136 272; do export RTB200_WAVE_MPATHS=$w; run --config c3 --spp 32 --pipeline wavefront; done
for w in 34 68 136; do export RTB200_WAVE_MPATHS=$w; run --config c4 --spp 128 --pipeline wavefront; done
export RTB200_WAVE_MPATHS=8; run --config c4 --spp 128 --pipeline regen
cat $L
