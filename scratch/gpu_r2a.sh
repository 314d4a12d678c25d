#!/bin/bash
# round 2, first GPU pass: the new tests, then the default bench line with its config legs
python -m pytest tests/test_gpu_round2.py -x -q -s 2>&1 | tail -40
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
tail -c 3000 gpurun_out/r2a_bench_n1.err
python scratch/show_bench.py gpurun_out/r2a_bench_n1.json
