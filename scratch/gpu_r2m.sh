#!/bin/bash
# round 2: sky colour in the primary cache + pool coop A/B on the interactive config; suite
python -m pytest tests -m gpu -q 2>&1 | tail -5
for coop in 0 1; do for rep in 1 2; do
RTB200_POOL_COOP=$coop python bench.py --config c5 --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 pool_coop=$coop p50', round(d['value'],4), 'p99', round(d['p99_ms'],4), 'render kernel', round(d['render_kernel_ms_p50'],4))"
done; done
python scratch/pool_sweep.py 2>&1 | grep -E "spp (1|2|4) "
