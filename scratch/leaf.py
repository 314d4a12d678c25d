import sys, numpy as np
sys.path.insert(0,"software-raytracer_b200/python"); import rtb200
from rtb200.scenes import synthetic_spheres, config3_camera
objs=np.load("tests/golden/bundled_scenes.npz")
big=synthetic_spheres(10000)
for name,o,cam,spp in [("Scene1",objs["Scene1"],rtb200.default_camera(),256),("Scene_indirect",objs["Scene_indirect"],rtb200.default_camera(),64),("C3-10k",big,config3_camera(rtb200.default_camera),16)]:
  for leaf in (1,2,4,8,16):
    t=rtb200.PathTracer(0); t.set_option(rtb200.RT_OPT_ACCEL,2); t.set_option(rtb200.RT_OPT_BVH_LEAF,leaf); t.set_scene(o); t.set_camera(cam)
    t.set_params(rtb200.default_params(width=1920,height=1080,mode=0,max_bounces=8)); t.reset_accumulation()
    t.render_spp(spp//4); t.sync(); best=1e9
    for i in range(2):
        t.reset_accumulation(); t.render_spp(spp); s=t.stats(); best=min(best,s.last_render_ms)
    print(name,"leaf",leaf,"ms %.2f"%best,"Gseg/s %.2f"%(s.segments/best/1e6)); t.close()
