#!/bin/bash
# r2z2: the chunk resolved as a whole by chunk_done (32 lanes, once per chunk) instead of per finished pixel; 32x1 pool tiles vs the previous build (tw8 = per-pixel resolve at the converged point, 8x4 tiles)
L=software-raytracer_b200/lib
python -m pytest tests/test_gpu_round2.py -q -x -k "render_frame" 2>&1 | tail -2
RTB200_LIB=$L/librt_b200_tw8.so python -m pytest tests/test_gpu_round2.py -q -x -k "render_frame" 2>&1 | tail -2
python scratch/ab_libs.py --reps 2 --cases c5f,c5f_1080,c5p,c5 $L/librt_b200_tw8.so $L/librt_b200.so 2>&1 | tee gpurun_out/r2z2_ab.txt
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
