#!/bin/bash
# r2ad: node steps per warp vote 1 / 2 / 3 / 4 in the persistent BVH kernels
L=software-raytracer_b200/lib
python scratch/ab_libs.py --reps 3 --cases c3w,c4w,c3s,c4s $L/librt_b200.so $L/librt_b200_u2.so $L/librt_b200_u3.so $L/librt_b200_u4.so 2>&1 | tee gpurun_out/r2ad_ab.txt
