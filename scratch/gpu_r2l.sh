#!/bin/bash
# round 2: single-instruction colour clamp: suite, bench line, launch list of the same command (flat back end pinned: under ncu AUTO's timing is perturbed)
python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err
tail -c 800 gpurun_out/r2l_bench_n1.err
python scratch/show_bench.py gpurun_out/r2l_bench_n1.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --accel flat"
$CMD > gpurun_out/r2l_plain.json 2> gpurun_out/r2l_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2l_launches.csv $CMD > gpurun_out/r2l_ncu_launches.log 2>&1
tail -1 gpurun_out/r2l_ncu_launches.log | cut -c1-200
