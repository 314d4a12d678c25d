#!/bin/bash
# r4t: ncu --set full capture of the megakernel (C2, 1024 spp) with the final build + the launch list of the default bench command
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --accel flat"
$CMD2 > gpurun_out/r4t_plain.json 2> gpurun_out/r4t_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_render_regen -s 4 -c 1 -o gpurun_out/r4t_regen_c2 -f $CMD2 > gpurun_out/r4t_ncu_regen.log 2>&1; tail -1 gpurun_out/r4t_ncu_regen.log | cut -c1-200
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --accel flat"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r4t_launches.csv $CMD > gpurun_out/r4t_ncu_launches.log 2>&1
tail -1 gpurun_out/r4t_ncu_launches.log | cut -c1-200
