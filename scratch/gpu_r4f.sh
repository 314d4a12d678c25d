#!/bin/bash
# r4f: why is the pixel-pool kernel slower than one pixel per lane at 64+ spp? one ncu capture of each at 64 spp
RTB200_POOL_COOP=1 python scratch/pool_vs_regen_ncu.py 64 > gpurun_out/r4f_plain.txt 2>&1 || exit 1
RTB200_POOL_COOP=1 ncu --set full --clock-control none --import-source on -k regex:k_render_ -s 2 -c 2 -f -o gpurun_out/r4f_pool_vs_regen python scratch/pool_vs_regen_ncu.py 64 > gpurun_out/r4f_ncu.log 2>&1
tail -3 gpurun_out/r4f_ncu.log; cat gpurun_out/r4f_plain.txt
