#!/bin/bash
cd $GRAFT_REPO_ROOT
for c in "c3 --spp 16" "c4 --spp 64"; do
 for r in 4 8 12 16; do for m in 4 8 12 16; do
  timeout 200 python bench.py --config $c --steps 2 --warmup 2 --no-cpu --pipeline wavefront --wf-refill $r --wf-node-min $m 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$c refill $r node_min $m value %.0f ms %.1f' % (d['value'], d['ms_per_step']))
"
 done; done
done
