#!/bin/bash
# r3b: why did the c4 leg of the default bench line take 44 ms per step (A/B: 22 ms)? tuner + wave sizes of the c4 leg alone and inside the full line
RTB200_DEBUG=1 python bench.py --config c4 --no-configs --no-cpu --steps 3 --warmup 3 2> gpurun_out/r3b_c4_alone.err | python scratch/show_bench.py /dev/stdin | head -2
grep "rtb200" gpurun_out/r3b_c4_alone.err | head -20
RTB200_DEBUG=1 python bench.py --no-cpu --steps 2 --warmup 3 2> gpurun_out/r3b_full.err > gpurun_out/r3b_full.json; python scratch/show_bench.py gpurun_out/r3b_full.json | cut -c1-200
grep "rtb200" gpurun_out/r3b_full.err | head -60
