#!/bin/bash
# r2u: per-octant cluster boxes + sign-bit masks in level 1, raw MUFU.RCP for the box tests, fused final add of the sphere cull
L=software-raytracer_b200/lib
python scratch/ab_libs.py --reps 2 --cases c2,c1,c5,c3w,c4w $L/librt_b200_prev.so $L/librt_b200.so 2>&1 | tee gpurun_out/r2u_ab.txt
python - <<'PY' 2>&1 | tee -a gpurun_out/r2u_ab.txt
# other bundled scenes (cube rooms): megakernel, 256 spp at 1080p
import sys, os, subprocess, json
for scene in ("Scene3", "Scene_indirect", "Scene2"):
    for lib in ("librt_b200_prev.so", "librt_b200.so"):
        env = dict(os.environ, RTB200_LIB=os.path.abspath("software-raytracer_b200/lib/" + lib))
        r = subprocess.run([sys.executable, "bench.py", "--scene", scene, "--spp", "256", "--steps", "3", "--warmup", "2", "--no-cpu", "--no-configs"], capture_output=True, text=True, env=env)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1]); print(scene, lib, round(d["ms_per_step"], 3), "ms", round(d["value"]), d["unit"], flush=True)
        except Exception as e:
            print(scene, lib, "FAILED", r.stderr[-300:])
PY
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
