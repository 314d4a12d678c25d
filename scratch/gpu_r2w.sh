#!/bin/bash
# r2w: rt_render_frame (render kernel resolves + streams the frame to the host surface) - tests, then 1-spp frame latency A/B
L=software-raytracer_b200/lib
python -m pytest tests/test_gpu_round2.py -q -x -k "render_frame" 2>&1 | tail -5
python scratch/ab_libs.py --reps 2 --cases c5,c5f,c5_1080,c5f_1080 $L/librt_b200.so 2>&1 | tee gpurun_out/r2w_ab.txt
RTB200_POOL_COOP=1 python scratch/ab_libs.py --reps 1 --cases c5,c5f,c5_1080,c5f_1080 $L/librt_b200.so 2>&1 | sed 's/^/POOL_COOP=1 /' | tee -a gpurun_out/r2w_ab.txt
python bench.py --config c5 --no-configs --no-cpu --steps 3 --warmup 3 > gpurun_out/r2w_bench_c5.json 2> gpurun_out/r2w_bench_c5.err; tail -c 1500 gpurun_out/r2w_bench_c5.json
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
