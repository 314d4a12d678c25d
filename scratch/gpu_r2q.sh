#!/bin/bash
# ncu: the bounce-round intersect kernel (first round of a wave = the big launch) on c3 and c4, and the shade kernel
for c in c3 c4; do
  CMD="python bench.py --config $c --no-configs --no-cpu --steps 1 --warmup 1 --pipeline wavefront"
  $CMD > gpurun_out/r2q_plain_$c.json 2> gpurun_out/r2q_plain_$c.err && \
  ncu --set full --clock-control none --import-source on -k regex:"k_wf_intersect_bvh|k_wf_shade" -s 16 -c 2 -o gpurun_out/r2q_wf_$c -f $CMD > gpurun_out/r2q_ncu_$c.log 2>&1
  tail -2 gpurun_out/r2q_ncu_$c.log
done
