# 6-CTA vs 7-CTA (STASH) form of the megakernel at mid-size grids: RTB200_STASH_MIN_TILES=1 forces the 7-CTA form, a huge value the 6-CTA one
import os, sys, hashlib, numpy as np
sys.path.insert(0, "software-raytracer_b200/python"); import rtb200
objs = np.load("tests/golden/bundled_scenes.npz")["Scene1"]
for (W, H), spps in (((1280, 720), (16, 64, 256)), ((960, 540), (64,)), ((640, 480), (64,)), ((2560, 1440), (64,))):
    t = rtb200.PathTracer(0); t.set_scene(objs); t.set_camera(rtb200.default_camera())
    t.set_params(rtb200.default_params(width=W, height=H, mode=0, max_bounces=8)); t.reset_accumulation()
    t.set_option(rtb200.RT_OPT_ACCEL, rtb200.RT_ACCEL_FLAT)
    for n in spps:
        for _ in range(2): t.render_spp(n)
        t.sync(); ms = []
        for _ in range(8):
            t.reset_accumulation(); t.render_spp(n); ms.append(t.stats().last_render_ms)
        a = t.read_accum()[0]
        print("min_tiles", os.environ.get("RTB200_STASH_MIN_TILES", "default"), W, H, "spp", n, "median %.3f min %.3f ms" % (np.median(ms), min(ms)), hashlib.sha256(a.tobytes()).hexdigest()[:12], flush=True)
    t.close()
