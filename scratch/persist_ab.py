# A/B of the persistent-grid form of the regeneration megakernel (RTB200_REGEN_PERSIST=<min spp>, 0 = off): ms per call + accumulation checksum
import os, sys, hashlib, numpy as np
sys.path.insert(0, "software-raytracer_b200/python"); import rtb200
scenes = np.load("tests/golden/bundled_scenes.npz")
for name, (W, H), spps in (("Scene1", (1920, 1080), (16, 64, 256, 1024)), ("Scene1", (640, 480), (64,)), ("Scene1", (1280, 720), (4, 8)), ("Scene_indirect", (1920, 1080), (64,)), ("Scene3", (1920, 1080), (64,))):
    t = rtb200.PathTracer(0); t.set_scene(scenes[name]); t.set_camera(rtb200.default_camera())
    t.set_params(rtb200.default_params(width=W, height=H, mode=0, max_bounces=8)); t.reset_accumulation()
    t.set_option(rtb200.RT_OPT_ACCEL, rtb200.RT_ACCEL_FLAT)
    for n in spps:
        for _ in range(2): t.render_spp(n)
        t.sync(); ms = []
        for _ in range(5 if n >= 1024 else 10):
            t.reset_accumulation(); t.render_spp(n); ms.append(t.stats().last_render_ms)
        a = t.read_accum()[0]
        print("persist", os.environ.get("RTB200_REGEN_PERSIST", "0"), name, W, H, "spp", n, "median %.3f min %.3f ms" % (np.median(ms), min(ms)), hashlib.sha256(a.tobytes()).hexdigest()[:12], flush=True)
    t.close()
