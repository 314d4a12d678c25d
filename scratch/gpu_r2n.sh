#!/bin/bash
# round 2: pixel-pool kernel with one scatter site
python -m pytest tests -m gpu -q 2>&1 | tail -5
for rep in 1 2; do
python bench.py --config c5 --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 p50', round(d['value'],4), 'p99', round(d['p99_ms'],4), 'render kernel', round(d['render_kernel_ms_p50'],4))"
done
python scratch/pool_sweep.py 2>&1 | grep -E "spp (1|2|4|8) "
