import sys, numpy as np
sys.path.insert(0,"software-raytracer_b200/python"); import rtb200
from rtb200.scenes import synthetic_spheres, config3_camera
objs=synthetic_spheres(10000)
for (w,h,spp) in [(1920,1080,16),(3840,2160,16)]:
  for accel in (2,1):
    if accel==1 and w==3840: continue
    t=rtb200.PathTracer(0); t.set_option(rtb200.RT_OPT_ACCEL,accel); t.set_scene(objs); t.set_camera(config3_camera(rtb200.default_camera))
    t.set_params(rtb200.default_params(width=w,height=h,mode=0,max_bounces=8)); t.reset_accumulation()
    n = spp if accel==2 else 1
    t.render_spp(n); t.sync(); t.reset_accumulation(); t.render_spp(n); s=t.stats()
    print("C3 10k spheres",w,h,"spp",n,"bvh" if accel==2 else "brute","ms %.2f"%s.last_render_ms,"Gseg/s %.3f"%(s.segments/s.last_render_ms/1e6),"Mpaths/s %.1f"%(s.paths/s.last_render_ms/1e3),"seg/path %.2f"%(s.segments/s.paths)); t.close()
