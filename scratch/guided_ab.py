# A/B of the guided pixel pool (RTB200_POOL_GUIDED=<min spp>, 0 = off -> one pixel per lane from 3 spp on): ms per call + accumulation checksum
import os, sys, hashlib, numpy as np
sys.path.insert(0, "software-raytracer_b200/python"); import rtb200
scenes = np.load("tests/golden/bundled_scenes.npz")
cases = (("Scene1", (1920, 1080), (4, 16, 64, 128, 256, 1024)), ("Scene1", (640, 480), (16, 64, 256)), ("Scene1", (1280, 720), (4, 8, 32)),
         ("Scene_indirect", (1920, 1080), (16, 64)), ("Scene3", (1920, 1080), (16, 64)), ("Scene2", (1920, 1080), (64,)))
for name, (W, H), spps in cases:
    t = rtb200.PathTracer(0); t.set_scene(scenes[name]); t.set_camera(rtb200.default_camera())
    t.set_params(rtb200.default_params(width=W, height=H, mode=0, max_bounces=8)); t.reset_accumulation()
    t.set_option(rtb200.RT_OPT_ACCEL, rtb200.RT_ACCEL_FLAT)
    for n in spps:
        for _ in range(2): t.render_spp(n)
        t.sync(); ms = []
        for _ in range(4 if n >= 1024 else 10):
            t.reset_accumulation(); t.render_spp(n); ms.append(t.stats().last_render_ms)
        a = t.read_accum()[0]
        print("guided", os.environ.get("RTB200_POOL_GUIDED", "0"), name, W, H, "spp", n, "median %.3f min %.3f ms" % (np.median(ms), min(ms)), hashlib.sha256(a.tobytes()).hexdigest()[:12], flush=True)
    t.close()
