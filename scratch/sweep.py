import sys, os, numpy as np
sys.path.insert(0,"software-raytracer_b200/python"); import rtb200
lib=os.environ.get("RTB_LIB")
if lib: rtb200._lib=rtb200.load_library(lib)
objs=np.load("tests/golden/bundled_scenes.npz")
for sc in ["Scene1","Scene_indirect"]:
  for accel in (1,2):
    t=rtb200.PathTracer(0); t.set_option(rtb200.RT_OPT_ACCEL,accel); t.set_scene(objs[sc]); t.set_camera(rtb200.default_camera())
    t.set_params(rtb200.default_params(width=1920,height=1080,mode=0,max_bounces=8)); t.reset_accumulation()
    t.render_spp(64); t.sync(); best=1e9
    for i in range(3):
        t.reset_accumulation(); t.render_spp(256); s=t.stats(); best=min(best,s.last_render_ms)
    print(lib or "default", sc, "brute" if accel==1 else "bvh", "ms %.2f"%best, "Gseg/s %.2f"%(s.segments/best/1e6)); t.close()
