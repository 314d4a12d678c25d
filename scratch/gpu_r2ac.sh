#!/bin/bash
# r2ac: persistent BVH kernels instantiated per node form (quantised / float); node steps per warp vote 1 vs 2
L=software-raytracer_b200/lib
python scratch/ab_libs.py --reps 3 --cases c3w,c4w,c3s,c4s $L/librt_b200_prev.so $L/librt_b200.so $L/librt_b200_u2.so 2>&1 | tee gpurun_out/r2ac_ab.txt
python -m pytest tests -m gpu -q -x -k "wavefront or stream or production or config3 or config4 or lane or mesh" 2>&1 | tail -2
