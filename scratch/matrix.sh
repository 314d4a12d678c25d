#!/bin/bash
# quick A/B matrix: accel x primary reuse, 1080p. usage: matrix.sh "scenes" "accels" "reuse flags" spp
cd $GRAFT_REPO_ROOT
SCENES=${1:-"Scene1 Scene_indirect"}; ACCELS=${2:-"brute bvh flat"}; SPP=${4:-256}
IFS=',' read -ra REUSE <<< "${3:-on,off}"
for scene in $SCENES; do
 for accel in $ACCELS; do
  for r in "${REUSE[@]}"; do
    flag=""; [ "$r" = "off" ] && flag="--no-primary-reuse"
    timeout 300 python bench.py --steps 2 --warmup 3 --spp $SPP --no-cpu --accel $accel $flag --scene $scene $EXTRA 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$scene $accel reuse=$r spp=$SPP value %.0f Mseg/s traced %.0f paths %.0f M/s ms/step %.2f seg/path %.3f' % (d['value'], d['traced_segments_per_s_M'], d['paths_per_s_M'], d['ms_per_step'], d['segments_per_path']))
"
  done
 done
done
