#!/bin/bash
# r2v: full ncu capture of the C2 megakernel (one 1024-spp launch) and of the 1-spp pixel-pool kernel at 720p, after the level-1 changes
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu --no-configs --accel flat"
$CMD2 > gpurun_out/r2v_plain2.json 2> gpurun_out/r2v_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:k_render_regen -s 4 -c 1 -o gpurun_out/r2v_regen_c2 -f $CMD2 > gpurun_out/r2v_ncu_regen.log 2>&1
tail -2 gpurun_out/r2v_ncu_regen.log | cut -c1-300
CMD3="python bench.py --config c5 --steps 1 --warmup 3 --no-cpu --no-configs"
$CMD3 > gpurun_out/r2v_plain3.json 2> gpurun_out/r2v_plain3.err && \
ncu --set full --clock-control none --import-source on -k regex:k_render_pool -s 40 -c 1 -o gpurun_out/r2v_pool_c5 -f $CMD3 > gpurun_out/r2v_ncu_pool.log 2>&1
tail -2 gpurun_out/r2v_ncu_pool.log | cut -c1-300
