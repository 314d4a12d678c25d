#!/usr/bin/env python
"""Generate tests/golden/* by running the REFERENCE'S OWN code (oracle/_ref/libref.so, built
from /root/reference by oracle/Makefile). Run here (CPU container, reference mounted); the
outputs are committed because /root/reference does not exist on the GPU box.

    python oracle/make_goldens.py            # ~5 min on 8 cores

TEST INFRASTRUCTURE ONLY.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from oracle_py import Oracle, Reference, build_ref, load_scene_json, REFERENCE_DIR, OBJECT_DTYPE  # noqa: E402

GOLD = os.path.normpath(os.path.join(HERE, "..", "tests", "golden"))
SCENES = ["Scene1", "Scene1_reflection", "Scene2", "Scene3", "Scene3_indirect", "Scene_indirect"]
SEED = (0x1234ABCD, 0x0BADC0DE)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def scene_path(name):
    return os.path.join(REFERENCE_DIR, "Scenes", name + ".json")


def rotated_camera(orc):
    """A non-trivial pose reached the way the viewer does it (Raytracer.cpp:394-395, 504-520)."""
    cam = orc.default_camera(70)
    orc.rotate_camera(cam, 0.35, [0, 1, 0])
    orc.rotate_camera(cam, -0.2, list(cam.right))
    cam.pos[0], cam.pos[1], cam.pos[2] = 1.5, 0.75, -2.0
    return cam


def main():
    build_ref()
    os.makedirs(GOLD, exist_ok=True)
    ref = Reference()
    orc = Oracle()
    meta = {"seed": list(SEED), "scenes": {}, "generated_by": "oracle/make_goldens.py via oracle/_ref (reference's own code)"}
    rng = np.random.default_rng(20261018)

    # 1. bundled scenes as numeric fixtures (+ hash of the reference's own Scene::Save output)
    scene_arrays = {}
    for s in SCENES:
        n = ref.load_scene(scene_path(s))
        objs = ref.objects()
        assert n == len(objs)
        assert objs.tobytes() == load_scene_json(scene_path(s)).tobytes()
        scene_arrays[s] = objs
        tmp = "/tmp/_ref_save_%s.json" % s
        ref.save_scene(tmp)
        with open(tmp, "rb") as f:
            saved = f.read()
        with open(scene_path(s), "rb") as f:
            original = f.read()
        with open(scene_path(s)) as f:
            doc = json.load(f)
        meta["scenes"][s] = {"n_objects": int(n), "save_sha256": hashlib.sha256(saved).hexdigest(),
                             "scene_name": doc["SceneName"], "names": [o["Name"] for o in doc["SceneObjects"]],
                             "save_bytes": len(saved), "save_equals_original_file": saved == original}
    np.savez_compressed(os.path.join(GOLD, "bundled_scenes.npz"), **scene_arrays)

    # 2. primary visibility AOVs: default camera and a rotated one
    aov = {}
    for s in SCENES:
        ref.load_scene(scene_path(s))
        for (w, h) in [(160, 120), (640, 480)] + ([(1920, 1080)] if s == "Scene1" else []):
            for cam_name, cam in (("default", orc.default_camera(55)), ("rotated", rotated_camera(orc))):
                if (w, h) == (1920, 1080) and cam_name != "default":
                    continue
                ref.setup(w, h, cam.fov_deg, 8, False, cam)
                ids, t, nrm, pt = ref.primary_aov()
                dirs = ref.ray_dirs()
                key = "%s_%dx%d_%s" % (s, w, h, cam_name)
                hit = ids >= 0
                meta.setdefault("aov", {})[key] = {
                    "hits": int(hit.sum()), "sum_t": float(t[hit].astype(np.float64).sum()),
                    "ids_sha256": sha(ids), "t_sha256": sha(t), "normal_sha256": sha(nrm), "point_sha256": sha(pt),
                    "dirs_sha256": sha(dirs)}
                if (w, h) != (1920, 1080):
                    aov[key + "_ids"] = ids.astype(np.int16)
                if (w, h) == (160, 120):
                    aov[key + "_t"] = t; aov[key + "_normal"] = nrm; aov[key + "_point"] = pt
    np.savez_compressed(os.path.join(GOLD, "primary_aov.npz"), **aov)
    cam = rotated_camera(orc)
    meta["rotated_camera"] = {"pos": list(cam.pos), "right": list(cam.right), "up": list(cam.up),
                              "forward": list(cam.forward), "fov_deg": cam.fov_deg}
    # Transform::RotateAboutAxis by the reference itself
    ref.setup(160, 120, 70, 8, False, orc.default_camera(70))
    ref.rotate_camera(0.35, [0, 1, 0])
    r = ref.camera()
    ref.rotate_camera(-0.2, list(r[1]))
    r = ref.camera()
    meta["rotated_camera_by_reference"] = {"right": r[1].tolist(), "up": r[2].tolist(), "forward": r[3].tolist()}

    # 3. arbitrary (secondary-like) rays through GetClosestObject, incl. the quirk cases
    rays = {}
    for s in ["Scene1", "Scene2", "Scene3", "Scene_indirect"]:
        ref.load_scene(scene_path(s))
        objs = scene_arrays[s]
        n = 4096
        org = np.zeros((n, 3), np.float32); dr = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        dr /= np.linalg.norm(dr, axis=1, keepdims=True).astype(np.float32)
        k = rng.integers(0, len(objs), n)
        # origins: on / just inside / just outside object surfaces, inside spheres, far away
        for i in range(n):
            o = objs[k[i]]
            u = rng.uniform(-1, 1, 3); u /= np.linalg.norm(u)
            ext = o["radius"] if o["type"] == 1 else float(np.max(o["half"]))
            scale = [1.0, 1.00001, 0.9999, 0.5, 0.0, 3.0][i % 6]
            org[i] = o["pos"] + (u * ext * scale).astype(np.float32)
        # exact zero direction components (the Box hole quirk) and axis-aligned rays
        dr[::17, 0] = 0; dr[5::29, 1] = 0; dr[7::31, 2] = 0
        dr[11::97] = np.array([0, 0, 1], np.float32)
        ids, t, nrm, pt = ref.trace_rays(org, dr)
        rays[s + "_org"] = org; rays[s + "_dir"] = dr; rays[s + "_id"] = ids.astype(np.int16)
        rays[s + "_t"] = t; rays[s + "_normal"] = nrm; rays[s + "_point"] = pt
    np.savez_compressed(os.path.join(GOLD, "trace_rays.npz"), **rays)

    # 4. environment colours
    d = rng.uniform(-1, 1, (4096, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    sun = -ref.env_constants()[:3]
    d[:64] = (sun + rng.normal(0, 0.08, (64, 3))).astype(np.float32)       # around the sun disc edge
    d[:64] /= np.linalg.norm(d[:64], axis=1, keepdims=True).astype(np.float32)
    d[64:96, 1] = 0                                                          # horizon
    d[96] = [0, 1, 0]; d[97] = [0, -1, 0]; d[98] = sun
    np.savez_compressed(os.path.join(GOLD, "env.npz"), dirs=d, rgb=ref.env_color(d), constants=ref.env_constants())

    # 5. per-sample radiance with rand() fed from the Philox stream (bit-comparable with the GPU)
    rad = {}
    for s in SCENES:
        ref.load_scene(scene_path(s))
        for cam_name, cam in (("default", orc.default_camera(55)), ("rotated", rotated_camera(orc))):
            w, h, nspp = 64, 48, 4
            for mb in ([8] if cam_name == "rotated" else [8, 2, 0]):
                ref.setup(w, h, cam.fov_deg, mb, False, cam)
                tot, per = ref.render_philox(SEED[0], SEED[1], 0, nspp, per_sample=True)
                key = "%s_%s_mb%d" % (s, cam_name, mb)
                rad[key + "_samples"] = per
                rad[key + "_sum"] = tot
        # a larger one by hash only, samples [8, 24)
        cam = orc.default_camera(55)
        ref.setup(160, 120, 55, 8, False, cam)
        tot, _ = ref.render_philox(SEED[0], SEED[1], 8, 16)
        meta.setdefault("radiance_sum_160x120_s8_n16", {})[s] = {"sha256": sha(tot), "mean_rgb": tot.mean(axis=(0, 1)).tolist()}
        rad[s + "_sum_160x120_s8_n16"] = tot
    np.savez_compressed(os.path.join(GOLD, "radiance_philox.npz"), **rad)

    # 6. preview (SIMPLEDRAW) shading incl. selection highlight + picking
    prev = {}
    for s in ["Scene1", "Scene2", "Scene3"]:
        ref.load_scene(scene_path(s))
        cam = rotated_camera(orc) if s == "Scene2" else orc.default_camera(55)
        ref.setup(160, 120, cam.fov_deg, 2, True, cam)
        sel = 64 if s != "Scene3" else 58
        ref.select(sel)
        tot, _ = ref.render_philox(0, 0, 0, 1)
        prev[s + "_selected%d" % sel] = tot
        ref.select(-1)
        tot, _ = ref.render_philox(0, 0, 0, 1)
        prev[s + "_noselect"] = tot
        picks = [(x, y, ref.pick(x, y)) for (x, y) in [(80, 60), (0, 0), (159, 119), (40, 100), (120, 30), (80, 61), (81, 60)]]
        meta.setdefault("pick_160x120", {})[s] = picks
    np.savez_compressed(os.path.join(GOLD, "preview.npz"), **prev)

    # 7. SetScreenPixel: running mean + Reinhard + ARGB8 pack on a value sweep
    w, h = 64, 16
    vals = np.concatenate([
        np.array([0, 1e-45, 1e-38, 1e-8, 1 / 255, 0.00392, 0.0039216, 0.5, 1, 2, 254 / 255, 255, 1e4, 1e30, 3e38,
                  np.inf, -1, -0.0, np.nan, 0.999999, 0.003921569, 127.5 / 255], np.float32),
        rng.uniform(0, 1, 200).astype(np.float32), rng.uniform(0, 600, 200).astype(np.float32),
        (10 ** rng.uniform(-6, 6, 200)).astype(np.float32)])
    rgba = np.zeros((h, w, 4), np.float32)
    flat = rgba.reshape(-1, 4)
    for c in range(3):
        flat[:, c] = rng.permutation(np.resize(vals, flat.shape[0]))
    ref.setup(w, h, 55, 8, False, None)
    surf, buf = ref.set_pixels(rgba, True, 1)
    res = {"rgba_in": rgba, "surface_setframe": surf, "buffer_setframe": buf}
    rgba2 = np.zeros_like(rgba); rgba2.reshape(-1, 4)[:, :3] = rng.uniform(0, 50, (flat.shape[0], 3)).astype(np.float32)
    surf2, buf2 = ref.set_pixels(rgba2, False, 7)          # running mean, 7th frame
    res.update({"rgba_in2": rgba2, "surface_mean7": surf2, "buffer_mean7": buf2})
    np.savez_compressed(os.path.join(GOLD, "resolve.npz"), **res)

    # 8. converged images by the reference with ITS OWN rand() (MSVC LCG, 16 strips): two independent
    #    runs give the reference's Monte Carlo noise floor for the statistical radiance test.
    conv = {}
    for s, spp in [("Scene1", 4096), ("Scene2", 2048), ("Scene_indirect", 1024)]:
        ref.load_scene(scene_path(s))
        ref.setup(160, 120, 55, 8, False, orc.default_camera(55))
        imgs = []
        for run in range(2):
            t0 = time.time()
            # run 0: MSVC semantics (all workers seeded 1); run 1: other seeds = an independent estimate
            sec, segs = ref.render_frames(spp, rng_mode=0, start_frame=1, count_segments=(run == 0),
                                          thread_seed=1 if run == 0 else 977)
            img = ref.color_buffer()[..., :3].copy()
            imgs.append(img)
            if run == 0:
                meta.setdefault("converged", {})[s] = {"spp": spp, "segments_per_path": segs / (160 * 120 * spp),
                                                         "cpu_seconds": sec}
            print(s, "run", run, "%.1fs" % (time.time() - t0), flush=True)
        conv[s + "_a"] = imgs[0]; conv[s + "_b"] = imgs[1]
        rmse = float(np.sqrt(np.mean((imgs[0] - imgs[1]) ** 2)))
        meta["converged"][s]["two_run_rmse_linear"] = rmse
        meta["converged"][s]["mean_rgb"] = imgs[0].mean(axis=(0, 1)).tolist()
    np.savez_compressed(os.path.join(GOLD, "converged_reference.npz"), **conv)

    with open(os.path.join(GOLD, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
