import sys, time, numpy as np
sys.path.insert(0, "software-raytracer_b200/python"); import rtb200
objs = np.load("tests/golden/bundled_scenes.npz")["Scene1"]
for (W, H) in ((1280, 720), (1920, 1080)):
    t = rtb200.PathTracer(0); t.set_scene(objs); t.set_camera(rtb200.default_camera())
    t.set_params(rtb200.default_params(width=W, height=H, mode=0, max_bounces=8)); t.reset_accumulation()
    for n in (1, 2, 4, 8, 16, 32):
        row = []
        for pool in (1, 2, 3, 4, 8):
            t.set_option(rtb200.RT_OPT_POOL_TILES, pool)
            for _ in range(5): t.render_spp(n)
            t.sync(); ms = []
            for _ in range(20):
                t.render_spp(n); ms.append(t.stats().last_render_ms)
            row.append("%d:%.3f" % (pool, np.median(ms)))
        print(W, H, "spp", n, " ".join(row))
    t.close()
