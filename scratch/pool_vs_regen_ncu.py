# one-pixel-per-lane megakernel vs the pixel-pool kernel at the same 64 spp (1080p Scene1): for an ncu capture of one launch each
import sys, numpy as np
sys.path.insert(0, "software-raytracer_b200/python"); import rtb200
objs = np.load("tests/golden/bundled_scenes.npz")["Scene1"]
t = rtb200.PathTracer(0); t.set_scene(objs); t.set_camera(rtb200.default_camera())
t.set_params(rtb200.default_params(width=1920, height=1080, mode=0, max_bounces=8)); t.reset_accumulation()
t.set_option(rtb200.RT_OPT_ACCEL, rtb200.RT_ACCEL_FLAT)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for pool in (1, 2):
    t.set_option(rtb200.RT_OPT_POOL_TILES, pool)
    for _ in range(3):
        t.render_spp(n)
    t.sync(); print(pool, t.stats().last_render_ms)
t.close()
