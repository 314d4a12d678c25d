#!/bin/bash
# r2ab: one retire point per pass, chunk resolve out of line (one call site)
L=software-raytracer_b200/lib
python -m pytest tests/test_gpu_round2.py -q -x -k "render_frame" 2>&1 | tail -2
python scratch/ab_libs.py --reps 2 --cases c5f,c5f_1080,c5p,c5 $L/librt_b200.so 2>&1 | tee gpurun_out/r2ab_ab.txt
