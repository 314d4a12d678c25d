#!/bin/bash
# round 2: leaf-ordered primitive slots: suite, pipelines on c3/c4, ncu of the stream kernel
python -m pytest tests -m gpu -q 2>&1 | tail -8
for p in regen wavefront stream; do for c in c3 c4; do
  python bench.py --config $c --no-configs --no-cpu --steps 3 --warmup 2 --pipeline $p 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c $p', round(d['value']), round(d['traced_segments_per_s_M']), round(d['ms_per_step'],2))"
done; done
for c in c3 c4; do
  CMD="python bench.py --config $c --no-configs --no-cpu --steps 1 --warmup 1 --pipeline stream"
  $CMD > gpurun_out/r2e_plain_$c.json 2> gpurun_out/r2e_plain_$c.err && \
  ncu --set full --clock-control none --import-source on -k regex:k_wf_stream -s 1 -c 1 -o gpurun_out/r2e_stream_$c -f $CMD > gpurun_out/r2e_ncu_$c.log 2>&1
  tail -2 gpurun_out/r2e_ncu_$c.log
done
