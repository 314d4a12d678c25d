# Does the persistent pixel-pool kernel (lanes pick up a new pixel the moment they finish one) beat one-pixel-per-lane at MANY samples
# per call? (scratch/pool_sweep.py only covers 1..32 spp.)  usage: RTB200_POOL_COOP=0|1 python scratch/pool_highspp.py
import os, sys, numpy as np
sys.path.insert(0, "software-raytracer_b200/python"); import rtb200
objs = np.load("tests/golden/bundled_scenes.npz")["Scene1"]
W, H = 1920, 1080
t = rtb200.PathTracer(0); t.set_scene(objs); t.set_camera(rtb200.default_camera())
t.set_params(rtb200.default_params(width=W, height=H, mode=0, max_bounces=8)); t.reset_accumulation()
t.set_option(rtb200.RT_OPT_ACCEL, rtb200.RT_ACCEL_FLAT)
ref = None
for n in (64, 256, 1024):
    row = []
    for pool in (1, 2, 4):
        t.set_option(rtb200.RT_OPT_POOL_TILES, pool)
        for _ in range(2): t.render_spp(n)
        t.sync(); ms = []
        for _ in range(4 if n == 1024 else 8):
            t.reset_accumulation(); t.render_spp(n); ms.append(t.stats().last_render_ms)
        a = t.read_accum()[0]
        if pool == 1: ref = a
        row.append("%d:%.3f ms%s" % (pool, np.median(ms), "" if pool == 1 else (" same" if np.array_equal(a.view(np.uint32), ref.view(np.uint32)) else " DIFF")))
    print("coop", os.environ.get("RTB200_POOL_COOP", "0"), "spp", n, " ".join(row), flush=True)
t.close()
