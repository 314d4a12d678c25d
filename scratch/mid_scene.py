import sys, numpy as np
sys.path.insert(0, "software-raytracer_b200/python"); import rtb200
from rtb200.scenes import synthetic_spheres, config3_camera
for n in (300, 1000, 3000):
    objs = synthetic_spheres(n)
    for pipe in (rtb200.RT_PIPELINE_REGEN, rtb200.RT_PIPELINE_WAVEFRONT):
        t = rtb200.PathTracer(0); t.set_option(rtb200.RT_OPT_PIPELINE, pipe)
        t.set_scene(objs); t.set_camera(config3_camera(rtb200.default_camera))
        t.set_params(rtb200.default_params(width=1920, height=1080, mode=0, max_bounces=8)); t.reset_accumulation()
        t.render_spp(8); t.sync(); best = 1e9
        for _ in range(3):
            t.reset_accumulation(); t.render_spp(32); s = t.stats(); best = min(best, s.last_render_ms)
        print(n, "pipeline", pipe, "accel", s.accel, "ms %.2f Gseg/s %.2f" % (best, s.segments / best / 1e6)); t.close()
