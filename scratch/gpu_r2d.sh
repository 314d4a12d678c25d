#!/bin/bash
# round 2: 256-bit node loads + dynamic pixel pools: suite, bench line, pool sweep
python -m pytest tests -m gpu -q 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
tail -c 1500 gpurun_out/r2d_bench_n1.err
python scratch/show_bench.py gpurun_out/r2d_bench_n1.json
for p in regen wavefront stream; do for c in c3 c4; do
  python bench.py --config $c --no-configs --no-cpu --steps 3 --warmup 2 --pipeline $p 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c $p', round(d['value']), round(d['traced_segments_per_s_M']), round(d['ms_per_step'],2))"
done; done
python scratch/pool_sweep.py 2>&1 | tail -14
