#!/bin/bash
# quick A/B matrix: accel x primary reuse on Scene1 and Scene_indirect, 1080p, 256 spp
cd $GRAFT_REPO_ROOT
for scene in Scene1 Scene_indirect; do
 for accel in brute bvh flat; do
  for reuse in "" "--no-primary-reuse"; do
    timeout 120 python bench.py --steps 2 --warmup 3 --spp 256 --no-cpu --accel $accel $reuse --scene $scene 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$scene $accel reuse=%s value %.0f Mseg/s traced %.0f paths %.0f M/s ms/step %.2f seg/path %.3f' % ('$reuse'=='', d['value'], d['traced_segments_per_s_M'], d['paths_per_s_M'], d['ms_per_step'], d['segments_per_path']))
"
  done
 done
done
