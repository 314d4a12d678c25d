"""A/B of library builds on one box (no torch): `python scratch/ab_libs.py [--reps R] [--cases c2,c3w,...] libA.so libB.so ...`.
Each (library, case) runs in its own process (RTB200_LIB) and prints the best and median device time of the render calls.
Cases: c1, c2 (megakernel), c5 / c5_1080 (1 spp: render + resolve into a host surface), c3w / c3s / c4w / c4s (wavefront / streaming)."""
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def child(case):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "software-raytracer_b200", "python"))
    import rtb200
    from rtb200.scenes import synthetic_spheres, config3_camera, heightfield_mesh, mesh_scene
    objs = np.load(os.path.join(ROOT, "tests", "golden", "bundled_scenes.npz"))["Scene1"]
    cam = rtb200.default_camera()
    mesh = None
    w, h, spp, pipe, reps = 1920, 1080, 1024, rtb200.RT_PIPELINE_AUTO, 5
    if case == "c1":
        w, h, spp, reps = 640, 480, 64, 20
    elif case in ("c5", "c5_1080", "c5f", "c5f_1080", "c5p"):    # c5f: one rt_render_frame per frame; c5p: the same into pageable memory
        w, h, spp, reps = (1280, 720, 1, 400) if case in ("c5", "c5f", "c5p") else (1920, 1080, 1, 400)
    elif case.startswith("c3"):
        objs, cam, w, h, spp = synthetic_spheres(10000), config3_camera(rtb200.default_camera), 3840, 2160, 16
    elif case.startswith("c4"):
        cam.pos[1] = 1.5; cam.pos[2] = -1.0
        objs, mesh, spp = mesh_scene(), heightfield_mesh(1024, 512), 64
    if case[2:] == "w": pipe = rtb200.RT_PIPELINE_WAVEFRONT
    if case[2:] == "s": pipe = rtb200.RT_PIPELINE_STREAM
    t = rtb200.PathTracer(0)
    t.set_option(rtb200.RT_OPT_PIPELINE, pipe)
    t.set_scene(objs)
    if mesh is not None:
        t.set_mesh(0, mesh[0], mesh[1])
    t.set_camera(cam)
    t.set_params(rtb200.default_params(width=w, height=h, mode=0, max_bounces=8, seed_lo=2026))
    t.reset_accumulation()
    out = {"case": case}
    if case.startswith("c5"):
        surf, _owner = rtb200.host_surface(w, h)
        if case == "c5p": surf = np.zeros((h, w), np.uint32)
        fused = case.startswith("c5f") or case == "c5p"
        for _ in range(50):
            if fused: t.render_frame(1, True, surf)
            else: t.render_spp(1); t.resolve_rgba8(True, surf)
        ms, wall = [], []
        for _ in range(reps):
            t0 = time.perf_counter()
            if fused: t.render_frame(1, True, surf)
            else: t.render_spp(1); t.resolve_rgba8(True, surf)
            wall.append((time.perf_counter() - t0) * 1e3)
        out["surface_sum"] = int(surf.astype(np.uint64).sum())
        for _ in range(100):
            t.render_spp(1); ms.append(t.stats().last_render_ms)
        if fused:
            fk = []
            for _ in range(100):
                t.render_frame(1, True, surf); fk.append(t.stats().last_render_ms)
            fk.sort(); out["frame_kernel_ms_p50"] = round(fk[len(fk) // 2], 4)
        ms.sort(); wall.sort()
        out.update(render_ms_p50=round(ms[len(ms) // 2], 4), frame_ms_p50=round(wall[len(wall) // 2], 4), frame_ms_p99=round(wall[int(len(wall) * 0.99)], 4))
    else:
        for _ in range(3):
            t.render_spp(spp)
        t.sync()
        ms = []
        s0 = t.stats()
        for _ in range(reps):
            t.render_spp(spp); ms.append(t.stats().last_render_ms)
        st = t.stats()
        ms.sort()
        out.update(ms_best=round(ms[0], 3), ms_med=round(ms[len(ms) // 2], 3), Gseg_s=round((st.total_segments - s0.total_segments) / reps / ms[len(ms) // 2] / 1e6, 3),
                   Gexec_s=round((st.total_traced_segments - s0.total_traced_segments) / reps / ms[len(ms) // 2] / 1e6, 3), pipeline=st.pipeline, accel=st.accel)
    acc, n = t.read_accum()
    out["checksum"] = "%016x" % (int(np.frombuffer(acc.tobytes(), dtype=np.uint64).sum(dtype=np.uint64)))
    out["samples"] = int(n)
    t.close()
    print(json.dumps(out), flush=True)


def main():
    args = sys.argv[1:]
    if args and args[0] == "--child":
        return child(args[1])
    reps, cases = 2, ["c2", "c1", "c5", "c5_1080", "c3w", "c3s", "c4w", "c4s"]
    libs = []
    i = 0
    while i < len(args):
        if args[i] == "--reps": reps = int(args[i + 1]); i += 2
        elif args[i] == "--cases": cases = args[i + 1].split(","); i += 2
        else: libs.append(args[i]); i += 1
    for rep in range(reps):
        for case in cases:
            for lib in libs:
                env = dict(os.environ, RTB200_LIB=os.path.abspath(lib))
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", case], capture_output=True, text=True, env=env)
                line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED rc=%d %s" % (r.returncode, r.stderr[-400:])
                print(os.path.basename(lib), line, flush=True)


if __name__ == "__main__":
    main()
