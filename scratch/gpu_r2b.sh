#!/bin/bash
# round 2: the whole GPU suite, then the default bench line
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
tail -c 2000 gpurun_out/r2b_bench_n1.err
python scratch/show_bench.py gpurun_out/r2b_bench_n1.json
