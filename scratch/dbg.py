import sys, numpy as np
sys.path.insert(0,"software-raytracer_b200/python"); sys.path.insert(0,"oracle"); sys.path.insert(0,"tests")
import rtb200
from oracle_py import Oracle, OrcCamera
objs=np.load("tests/golden/bundled_scenes.npz")["Scene1"]
w,h=64,48
import os
if os.environ.get("RTB_LIB"): rtb200._lib=rtb200.load_library(os.environ["RTB_LIB"])
t=rtb200.PathTracer(0); t.set_scene(objs); t.set_camera(rtb200.default_camera())
t.set_params(rtb200.default_params(width=w,height=h,mode=0,max_bounces=8,seed_lo=0x1234ABCD,seed_hi=0x0BADC0DE)); t.reset_accumulation()
t.render_spp(1); acc,_=t.read_accum()
orc=Oracle(); cam=OrcCamera(); cam.right[0]=1; cam.up[1]=1; cam.forward[2]=1; cam.fov_deg=55
p=orc.default_params(width=w,height=h,max_bounces=8,mode=0,seed_lo=0x1234ABCD,seed_hi=0x0BADC0DE)
want,_,segs=orc.render(objs,cam,p,0,1)
ids=t.read_aov()[0]
d=np.abs(acc[...,:3]-want)/np.maximum(np.abs(want),1e-6)
bad=np.argwhere(d.max(axis=2)>1e-5)
print("bad pixels",len(bad),"of",w*h, "gpu segs",t.stats().segments,"oracle",segs)
for y,x in bad[:20]: print(y,x,ids[y,x],acc[y,x,:3],want[y,x], d[y,x].max())
for mb in (0,1,2):
    t.set_params(max_bounces=mb); t.reset_accumulation(); t.render_spp(1); a,_=t.read_accum()
    p.max_bounces=mb; wnt,_,sg=orc.render(objs,cam,p,0,1)
    dd=np.abs(a[...,:3]-wnt)/np.maximum(np.abs(wnt),1e-6)
    print("mb",mb,"bad",(dd.max(axis=2)>1e-5).sum(), t.stats().segments, sg)
