#!/bin/bash
# A/B: streaming kernel at 8 (64 regs, spills) / 7 / 6 CTAs per SM
for lib in librt_b200.so librt_b200_s7.so librt_b200_s6.so; do for c in c3 c4; do
  RTB200_LIB=software-raytracer_b200/lib/$lib python bench.py --config $c --no-configs --no-cpu --steps 3 --warmup 2 --pipeline stream 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib $c stream', round(d['value']), round(d['traced_segments_per_s_M']), round(d['ms_per_step'],2))"
done; done
