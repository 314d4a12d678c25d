#!/bin/bash
# r2y: pool tile shape (8x4 / 16x2 / 32x1: 64 / 128 / 128-byte rows per chunk store) for the fused frame; pageable destination isolates the PCIe stores
L=software-raytracer_b200/lib
python scratch/ab_libs.py --reps 2 --cases c5f,c5f_1080,c5p,c5 $L/librt_b200.so $L/librt_b200_tw16.so $L/librt_b200_tw32.so 2>&1 | tee gpurun_out/r2y_ab.txt
RTB200_LIB=$L/librt_b200_tw16.so python -m pytest tests/test_gpu_round2.py -q -x -k "render_frame" 2>&1 | tail -2
RTB200_LIB=$L/librt_b200_tw32.so python -m pytest tests/test_gpu_round2.py -q -x -k "render_frame" 2>&1 | tail -2
