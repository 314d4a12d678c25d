#!/bin/bash
# r2t: shared-reciprocal normalise + Philox round keys as constants + out-of-line streaming shade, against the previous build
L=software-raytracer_b200/lib
python - <<'PY'
import sys
sys.path.insert(0, "software-raytracer_b200/python")
import rtb200
t = rtb200.PathTracer(0)
print("selftest uniform", t.selftest(0), "normalize", t.selftest(1), flush=True)
t.close()
PY
python scratch/ab_libs.py --reps 2 $L/librt_b200_base.so $L/librt_b200.so 2>&1 | tee gpurun_out/r2t_ab.txt
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
