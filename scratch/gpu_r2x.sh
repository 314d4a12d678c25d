#!/bin/bash
# r2x: checkpoint after the round-2b kernel work: suite, both bench arms as the driver calls them, launch list, c5 A/B of the cursor reset
L=software-raytracer_b200/lib
python -c "import __graft_entry__ as g; g.smoke()"
python -m pytest tests -m gpu -q 2>&1 | tail -3
python scratch/ab_libs.py --reps 2 --cases c5f,c5f_1080,c5 $L/librt_b200.so 2>&1 | tee gpurun_out/r2x_ab.txt
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2x_bench_reference_arm.json 2> gpurun_out/r2x_bench_reference_arm.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2x_bench_n1.json 2> gpurun_out/r2x_bench_n1.err
tail -c 400 gpurun_out/r2x_bench_n1.err
python scratch/show_bench.py gpurun_out/r2x_bench_n1.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --accel flat"
$CMD > gpurun_out/r2x_plain.json 2> gpurun_out/r2x_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2x_launches.csv $CMD > gpurun_out/r2x_ncu_launches.log 2>&1
tail -1 gpurun_out/r2x_ncu_launches.log | cut -c1-200
