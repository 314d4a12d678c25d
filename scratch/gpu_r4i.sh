#!/bin/bash
# r4i: ncu capture of the guided pixel pool over the live-pixel list at 64 spp (1080p Scene1), third launch
cat > /tmp/g.py <<'P'
import sys, numpy as np
sys.path.insert(0, "software-raytracer_b200/python"); import rtb200
objs = np.load("tests/golden/bundled_scenes.npz")["Scene1"]
t = rtb200.PathTracer(0); t.set_scene(objs); t.set_camera(rtb200.default_camera())
t.set_params(rtb200.default_params(width=1920, height=1080, mode=0, max_bounces=8)); t.reset_accumulation()
t.set_option(rtb200.RT_OPT_ACCEL, rtb200.RT_ACCEL_FLAT)
for _ in range(3): t.render_spp(64)
t.sync(); print(t.stats().last_render_ms); t.close()
P
RTB200_POOL_GUIDED=3 python /tmp/g.py > gpurun_out/r4i_plain.txt 2>&1 || exit 1
RTB200_POOL_GUIDED=3 ncu --set full --clock-control none --import-source on -k regex:k_render_pool -s 2 -c 1 -f -o gpurun_out/r4i_guided python /tmp/g.py > gpurun_out/r4i_ncu.log 2>&1
tail -2 gpurun_out/r4i_ncu.log; cat gpurun_out/r4i_plain.txt
