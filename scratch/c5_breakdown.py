import sys, time, numpy as np
sys.path.insert(0, "software-raytracer_b200/python"); import rtb200
import torch
objs = np.load("tests/golden/bundled_scenes.npz")["Scene1"]
W, H = 1280, 720
t = rtb200.PathTracer(0); t.set_scene(objs); t.set_camera(rtb200.default_camera())
t.set_params(rtb200.default_params(width=W, height=H, mode=0, max_bounces=8)); t.reset_accumulation()
out = np.zeros((H, W), np.uint32)
pinned = torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
for name, buf in (("pageable", out), ("pinned", pinned)):
    for _ in range(20): t.render_spp(1); t.resolve_rgba8(True, buf)
    n = 300; a = []; r = []; rs = []
    for _ in range(n):
        t0 = time.perf_counter(); t.render_spp(1); t1 = time.perf_counter(); t.resolve_rgba8(True, buf); t2 = time.perf_counter()
        a.append((t2 - t0) * 1e3); r.append((t1 - t0) * 1e3); rs.append((t2 - t1) * 1e3)
    st = t.stats()
    print(name, "frame p50 %.3f ms  render call %.3f  resolve call %.3f  | device render %.3f resolve %.3f" % (np.median(a), np.median(r), np.median(rs), st.last_render_ms, st.last_resolve_ms))
