#!/bin/bash
# r3c: legs sampled outside their timed steps: variance of the c3 / c4 legs over three full lines, then both arms as the driver calls them
for i in 1 2 3; do
python bench.py --no-cpu --steps 3 --warmup 3 2> /dev/null > gpurun_out/r3c_full_$i.json; python scratch/show_bench.py gpurun_out/r3c_full_$i.json | grep -E "'c3'|'c4'|'c5'" | cut -c1-150
done
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r3c_bench_reference_arm.json 2> gpurun_out/r3c_bench_reference_arm.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r3c_bench_n1.json 2> gpurun_out/r3c_bench_n1.err
python scratch/show_bench.py gpurun_out/r3c_bench_n1.json | cut -c1-220
