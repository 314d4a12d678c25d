"""Sweep of RT_OPT_WF_REFILL x RT_OPT_WF_NODE_MIN for the bounce-round wavefront pipeline on c3 / c4 (one context per scene)."""
import sys
import numpy as np
sys.path.insert(0, "software-raytracer_b200/python")
import rtb200
from rtb200.scenes import synthetic_spheres, config3_camera, heightfield_mesh, mesh_scene


def run(name, objs, cam, w, h, spp, mesh=None):
    t = rtb200.PathTracer(0)
    t.set_option(rtb200.RT_OPT_PIPELINE, rtb200.RT_PIPELINE_WAVEFRONT)
    t.set_scene(objs)
    if mesh is not None:
        t.set_mesh(0, mesh[0], mesh[1])
    t.set_camera(cam)
    t.set_params(rtb200.default_params(width=w, height=h, mode=0, max_bounces=8, seed_lo=2026))
    t.reset_accumulation()
    for refill in (4, 8, 12, 16, 24):
        row = []
        for node_min in (4, 8, 12, 16, 24):
            t.set_option(rtb200.RT_OPT_WF_REFILL, refill); t.set_option(rtb200.RT_OPT_WF_NODE_MIN, node_min)
            t.render_spp(spp); t.sync()
            ms = []
            for _ in range(3):
                t.render_spp(spp); ms.append(t.stats().last_render_ms)
            row.append("%d:%.2f" % (node_min, min(ms)))
        print(name, "refill", refill, "node_min ->", " ".join(row), flush=True)
    t.close()


run("c3", synthetic_spheres(10000), config3_camera(rtb200.default_camera), 3840, 2160, 16)
cam = rtb200.default_camera(); cam.pos[1] = 1.5; cam.pos[2] = -1.0
run("c4", mesh_scene(), cam, 1920, 1080, 64, heightfield_mesh(1024, 512))
