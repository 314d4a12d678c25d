#!/bin/bash
# round 2 profile set: launch list of the default bench command, full capture of the C2 megakernel, converged C2 test output
python -m pytest tests/test_gpu_round2.py -q -s -k "config2_converged or against_float64" 2>&1 | grep -E "C2 1920|mesh vs|passed|failed"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/r2j_plain.json 2> gpurun_out/r2j_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2j_launches.csv $CMD > gpurun_out/r2j_ncu_launches.log 2>&1
tail -2 gpurun_out/r2j_ncu_launches.log | cut -c1-300
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu --no-configs --accel flat"
$CMD2 > gpurun_out/r2j_plain2.json 2> gpurun_out/r2j_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:k_render_regen -s 4 -c 1 -o gpurun_out/r2j_regen_c2 -f $CMD2 > gpurun_out/r2j_ncu_regen.log 2>&1
tail -2 gpurun_out/r2j_ncu_regen.log | cut -c1-300
