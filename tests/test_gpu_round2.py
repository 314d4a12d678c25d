"""GPU (B200), second round: the persistent primary-hit cache, device traversal counters, the production-shape
configurations of BASELINE.json (C2 converged at 1920x1080 x 1024 spp against the reference's own renders, C3 at
3840x2160 and C4 at 1080p through RT_PIPELINE_AUTO / the wavefront pipeline), the independent triangle check, option
validation, the library-owned multi-GPU group and the C++ host driver (bin/rt_headless) run on the device."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import rtb200
from conftest import GOLD, ROOT, SEED, make_camera
from rtb200.scenes import synthetic_spheres, config3_camera, heightfield_mesh, mesh_scene
from test_gpu_parity import bits, close, setup, _mesh_cam

pytestmark = pytest.mark.gpu


# ---- primary-hit cache kept across rt_render_spp calls ------------------------------------------------------------
@pytest.mark.parametrize("scene", ["Scene1", "Scene3_indirect"])
def test_primary_cache_persists_across_calls_and_is_invalidated(tracer, scenes, meta, scene):
    w, h = 333, 77
    npx = w * h
    cam_b = make_camera(rtb200.RtCamera, meta, rotated=True)
    out = {}
    try:
        for reuse in (1, 0):
            tracer.set_option(rtb200.RT_OPT_PRIMARY_REUSE, reuse)
            setup(tracer, scenes[scene], w, h)
            traced = []
            for n in (1, 1, 3, 20):                      # a viewer's loop: 1 spp per call, static camera (Raytracer.cpp:572-590)
                tracer.render_spp(n)
                traced.append(tracer.stats().traced_segments)
            st = tracer.stats()
            a = tracer.read_accum()[0]
            tracer.set_camera(cam_b)                     # a camera move invalidates the cache; the scene does too
            tracer.reset_accumulation()
            tracer.render_spp(2)
            t_move = tracer.stats().traced_segments
            b = tracer.read_accum()[0]
            tracer.set_scene(scenes[scene][:-1])
            tracer.render_spp(2)
            c = tracer.read_accum()[0]
            out[reuse] = (a, b, c, st.segments, traced, t_move, st.paths)
    finally:
        tracer.set_option(rtb200.RT_OPT_PRIMARY_REUSE, 1)
    on, off = out[1], out[0]
    for k in range(3):                                   # identical bits with and without the cache, before and after invalidation
        assert np.array_equal(bits(on[k]), bits(off[k])), k
    assert on[3] == off[3] and on[6] == off[6] == npx * 25          # segments DELIVERED and paths are the same
    # executed queries: every pixel's primary ray exactly once for the four calls, and once more after the camera move
    assert off[4][-1] == off[3]
    assert on[4][-1] == off[4][-1] - npx * 25 + npx
    secondary_2spp = off[5] - npx * 2
    assert on[5] == secondary_2spp + npx


def test_same_camera_resubmitted_keeps_the_cache(tracer, scenes):
    setup(tracer, scenes["Scene1"], 160, 120)
    tracer.render_spp(1)
    t1 = tracer.stats().total_traced_segments
    tracer.set_camera(rtb200.default_camera())          # a host that sets its (unchanged) camera every frame
    tracer.render_spp(1)
    t2 = tracer.stats().total_traced_segments
    tracer.set_params(tracer.params)                      # same parameters again
    tracer.render_spp(1)
    t3 = tracer.stats().total_traced_segments
    assert t2 - t1 < 160 * 120 * 1.6 and abs((t3 - t2) - (t2 - t1)) < 0.1 * (t2 - t1)    # no second primary pass (it would add 160 * 120)


def test_environment_change_drops_the_cached_sky_colours(tracer, scenes):
    """The cache keeps GetEnvironmentColor(d) for pixels whose primary ray misses: new sky / sun parameters must invalidate it."""
    out = {}
    try:
        for reuse in (1, 0):
            tracer.set_option(rtb200.RT_OPT_PRIMARY_REUSE, reuse)
            setup(tracer, scenes["Scene1"], 160, 120)
            tracer.render_spp(2)
            a = tracer.read_accum()[0]
            tracer.params.sky[0], tracer.params.sky[1], tracer.params.sky[2] = 9.0, 1.0, 0.5
            tracer.params.sun_dir[0], tracer.params.sun_dir[1], tracer.params.sun_dir[2] = 0.0, -0.6, -0.8
            tracer.set_params(tracer.params)
            tracer.reset_accumulation(); tracer.render_spp(2)
            out[reuse] = (a, tracer.read_accum()[0])
    finally:
        tracer.set_option(rtb200.RT_OPT_PRIMARY_REUSE, 1)
    assert np.array_equal(bits(out[1][0]), bits(out[0][0])) and np.array_equal(bits(out[1][1]), bits(out[0][1]))
    assert not np.array_equal(out[1][0], out[1][1])


# ---- traversal counters (bytes per segment from node visits and primitive tests) -----------------------
def test_traversal_counters_equal_the_host_emulation(tracer):
    """The counting instantiation of the per-ray BVH loop must visit exactly the nodes and test exactly the primitives the
    same code does when compiled for the host (tests/host_emu), where the counters were first defined."""
    from test_device_logic_cpu import build_emu
    lib = C.CDLL(build_emu())
    lib.emu_render.restype = C.c_longlong
    objs = synthetic_spheres(300, seed=4, cubes_every=11)
    w, h, n = 96, 64, 3
    cam = config3_camera(rtb200.default_camera)
    par = rtb200.default_params(width=w, height=h, mode=rtb200.RT_MODE_PATH, max_bounces=8, seed_lo=SEED[0], seed_hi=SEED[1])
    stats = (C.c_longlong * 6)()
    lib.emu_bvh_stats(stats)
    out = np.zeros((h, w, 3), np.float32)
    o = np.ascontiguousarray(objs, rtb200.OBJECT_DTYPE)
    segs = lib.emu_render(o.ctypes.data_as(C.c_void_p), len(o), C.byref(cam), C.byref(par), 1, C.c_uint32(0), n,
                          out.ctypes.data_as(C.c_void_p), None, None, None, None)
    lib.emu_bvh_stats(stats)
    emu_nodes, emu_prims = stats[2], stats[5]
    try:
        tracer.set_option(rtb200.RT_OPT_ACCEL, rtb200.RT_ACCEL_BVH)
        tracer.set_option(rtb200.RT_OPT_PIPELINE, rtb200.RT_PIPELINE_REGEN)
        tracer.set_option(rtb200.RT_OPT_PRIMARY_REUSE, 0)                 # the emulation re-traces every primary ray
        tracer.set_option(rtb200.RT_OPT_TRAVERSAL_STATS, 1)
        setup(tracer, objs, w, h, cam)
        tracer.render_spp(n)
        ts, st = tracer.traversal_stats(), tracer.stats()
        got = tracer.read_accum()[0]
        # with the counters off nothing is counted and the image is the same
        tracer.set_option(rtb200.RT_OPT_TRAVERSAL_STATS, 0)
        tracer.reset_accumulation(); tracer.render_spp(n)
        assert tracer.traversal_stats().node_visits == 0
        assert np.array_equal(bits(got), bits(tracer.read_accum()[0]))
    finally:
        for opt, v in ((rtb200.RT_OPT_ACCEL, 0), (rtb200.RT_OPT_PIPELINE, 0), (rtb200.RT_OPT_PRIMARY_REUSE, 1), (rtb200.RT_OPT_TRAVERSAL_STATS, 0)):
            tracer.set_option(opt, v)
    assert st.segments == segs == ts.queries
    assert ts.node_visits == emu_nodes and ts.prim_tests == emu_prims and ts.node_bytes == 64
    assert ts.sphere_tests + ts.cube_tests == ts.prim_tests and ts.cube_tests > 0 and ts.tri_tests == 0
    assert close(got[..., :3], out)


# ---- BASELINE configs at production shape ------------------------------------------------------------------------
def _auto_vs_megakernel(tracer, setup_scene, w, h, spp, small_wave_mpaths):
    """RT_PIPELINE_AUTO (must pick one of the two persistent-thread pipelines) vs the megakernel vs the bounce-round wavefront
    pipeline vs the streaming kernel, the last two also forced into many small waves: bit-identical accumulation buffers and
    identical segment counts; traversal counters sane."""
    out = {}
    try:
        for name, pipe, mp, cnt in (("auto", rtb200.RT_PIPELINE_AUTO, 0, 0), ("regen", rtb200.RT_PIPELINE_REGEN, 0, 0),
                                    ("wavefront", rtb200.RT_PIPELINE_WAVEFRONT, 0, 0), ("stream", rtb200.RT_PIPELINE_STREAM, 0, 0),
                                    ("stream_waves", rtb200.RT_PIPELINE_STREAM, small_wave_mpaths, 1),
                                    ("waves", rtb200.RT_PIPELINE_WAVEFRONT, small_wave_mpaths, 1)):
            tracer.set_option(rtb200.RT_OPT_PIPELINE, pipe)
            tracer.set_option(rtb200.RT_OPT_WF_WAVE_MPATHS, mp)
            tracer.set_option(rtb200.RT_OPT_TRAVERSAL_STATS, cnt)
            setup_scene()
            tracer.render_spp(spp)
            st = tracer.stats()
            out[name] = (tracer.read_accum()[0], st.segments, st.traced_segments, st.pipeline, st.accel, st.last_render_ms, tracer.traversal_stats())
    finally:
        tracer.set_option(rtb200.RT_OPT_PIPELINE, rtb200.RT_PIPELINE_AUTO)
        tracer.set_option(rtb200.RT_OPT_WF_WAVE_MPATHS, 0)
        tracer.set_option(rtb200.RT_OPT_TRAVERSAL_STATS, 0)
    assert out["auto"][3] in (rtb200.RT_PIPELINE_WAVEFRONT, rtb200.RT_PIPELINE_STREAM) and out["auto"][4] == rtb200.RT_ACCEL_BVH
    assert out["regen"][3] == rtb200.RT_PIPELINE_REGEN and out["waves"][3] == out["wavefront"][3] == rtb200.RT_PIPELINE_WAVEFRONT
    assert out["stream"][3] == out["stream_waves"][3] == rtb200.RT_PIPELINE_STREAM
    for name in ("regen", "wavefront", "stream", "stream_waves", "waves"):
        assert np.array_equal(bits(out["auto"][0]), bits(out[name][0])), name
        assert out["auto"][1:3] == out[name][1:3], name
    for name in ("stream_waves", "waves"):
        ts = out[name][6]
        assert ts.queries == out[name][2] - w * h, name      # every executed query but the cached primaries went through the counting kernel
    a = out["auto"][0]
    assert np.all(np.isfinite(a)) and a[..., :3].max() > 0
    ts = out["waves"][6]
    assert 5 < ts.node_visits / ts.queries < 400 and 0.2 < ts.prim_tests / ts.queries < 60
    print("%dx%d x %d spp: auto(pipeline %d) %.1f ms, megakernel %.1f ms, wavefront %.1f ms, stream %.1f ms; small waves: wavefront %.1f ms, stream %.1f ms; %.1f node visits, %.1f primitive tests per query"
          % (w, h, spp, out["auto"][3], out["auto"][5], out["regen"][5], out["wavefront"][5], out["stream"][5], out["waves"][5], out["stream_waves"][5],
             ts.node_visits / ts.queries, ts.prim_tests / ts.queries))
    return out


def test_config3_10k_spheres_4k_through_auto_pipeline(tracer):
    """BASELINE.json configs[2] at its stated shape: 10 000 random spheres, 3840x2160, the pipeline RT_PIPELINE_AUTO selects."""
    objs = synthetic_spheres(10000)
    cam = config3_camera(rtb200.default_camera)
    w, h = 3840, 2160
    out = _auto_vs_megakernel(tracer, lambda: setup(tracer, objs, w, h, cam), w, h, 16, 24)   # 24 Mi paths: 2 samples per wave, 8 waves
    assert 2.0 < out["auto"][1] / (w * h * 16) < 6.0
    tracer.set_scene(objs[:1])                            # drop the big buffers' scene before the next test


def test_config4_one_million_triangles_1080p_through_auto_pipeline(tracer):
    """BASELINE.json configs[3]: 1 048 576-triangle mesh, 1920x1080, RT_PIPELINE_AUTO (wavefront) vs megakernel vs small waves."""
    v, tr = heightfield_mesh(1024, 512)
    objs = mesh_scene()

    def scene():
        setup(tracer, objs, 1920, 1080, _mesh_cam(rtb200.RtCamera))
        tracer.set_mesh(0, v, tr)
    out = _auto_vs_megakernel(tracer, scene, 1920, 1080, 16, 6)      # 6 Mi paths: 3 samples per wave
    assert out["waves"][6].tri_tests > 0
    tracer.set_scene(objs[1:])


def test_config2_converged_image_at_1080p_1024spp_vs_reference(tracer, scenes, meta):
    """BASELINE.json configs[1] at its own size: Scene1 1920x1080, 1024 spp, depth 8, against TWO renders of the same frame
    by the reference's own code with its own rand() (oracle/make_goldens_c2.py): the GPU image must be as close to each
    reference run as they are to each other (RMSE <= 1.1 x their two-run noise floor). PSNR of the tonemapped images stated."""
    path = os.path.join(GOLD, "converged_c2_1080p.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/converged_c2_1080p.npz not generated (oracle/make_goldens_c2.py)")
    z = np.load(path)
    ref_a, ref_b = z["a"].astype(np.float32), z["b"].astype(np.float32)
    m = meta["converged_c2"]
    setup(tracer, scenes["Scene1"], m["width"], m["height"])
    tracer.render_spp(m["spp"])
    acc, n = tracer.read_accum()
    st = tracer.stats()
    img = acc[..., :3] / np.float32(n)
    rmse = lambda x, y: float(np.sqrt(np.mean((x.astype(np.float64) - y) ** 2)))
    floor = rmse(ref_a, ref_b)
    ga, gb = rmse(img, ref_a), rmse(img, ref_b)
    tm = lambda x: x / (1 + x)
    psnr_a, psnr_ref = -20 * np.log10(rmse(tm(img), tm(ref_a))), -20 * np.log10(rmse(tm(ref_a), tm(ref_b)))
    print("C2 1920x1080 x %d spp: rmse gpu-refA %.4f gpu-refB %.4f floor(refA-refB) %.4f (meta %.4f); tonemapped PSNR gpu-refA %.2f dB, refA-refB %.2f dB; %.1f ms"
          % (n, ga, gb, floor, m["two_run_rmse_linear"], psnr_a, psnr_ref, st.last_render_ms))
    assert n == m["spp"] and abs(floor - m["two_run_rmse_linear"]) < 0.02 * floor       # float16 storage does not move the floor
    assert ga <= 1.10 * floor and gb <= 1.10 * floor
    assert psnr_a >= psnr_ref - 0.5
    assert np.allclose(img.mean(axis=(0, 1)), ref_a.mean(axis=(0, 1)), rtol=0.01)
    assert abs(st.segments / st.paths - m["segments_per_path"]) < 0.01
    # sky pixels carry no noise: they must match the reference's running mean to float rounding
    sky = tracer.read_aov()[0] < 0
    assert sky.mean() > 0.3 and np.allclose(img[sky], ref_a[sky], rtol=2e-3)           # float16 + running-mean rounding


# ---- mesh extension pinned from outside the product's own formula ---------------------------------------------------
def test_mesh_against_float64_moller_trumbore(tracer):
    from mt_reference import moller_trumbore, check_against_mt
    from test_mesh_independent_cpu import sampled_rays
    v, tr = heightfield_mesh(128, 128, seed=17)          # 32 768 triangles
    objs = mesh_scene()[:1]
    org, d = sampled_rays(2048, 3)
    res = {}
    try:
        for accel in (rtb200.RT_ACCEL_BVH, rtb200.RT_ACCEL_BRUTE):
            tracer.set_option(rtb200.RT_OPT_ACCEL, accel)
            setup(tracer, objs, 64, 48, _mesh_cam(rtb200.RtCamera))
            tracer.set_mesh(0, v, tr)
            res[accel] = tracer.trace_rays(org, d)
    finally:
        tracer.set_option(rtb200.RT_OPT_ACCEL, rtb200.RT_ACCEL_AUTO)
        tracer.set_scene(mesh_scene()[1:])
    ids, t, nrm, pt = res[rtb200.RT_ACCEL_BVH]
    for x, y in zip(res[rtb200.RT_ACCEL_BVH], res[rtb200.RT_ACCEL_BRUTE]):
        assert np.array_equal(bits(x), bits(y))
    mt = moller_trumbore(v.astype(np.float64) + objs["pos"][0].astype(np.float64), tr, org, d)
    n_hit, n_diff = check_against_mt(ids, t, nrm, mt)
    print("mesh vs float64 Moeller-Trumbore: %d of %d rays hit, %d hit/miss differences (all at triangle edges)" % (n_hit, len(org), n_diff))
    assert n_hit > 700 and n_diff <= 4


# ---- option validation ---------------------------------------------------------------------------
def test_set_option_rejects_out_of_range_values(tracer):
    bad = [(rtb200.RT_OPT_ACCEL, 7), (rtb200.RT_OPT_ACCEL, -1), (rtb200.RT_OPT_PIPELINE, 4), (rtb200.RT_OPT_BVH_THRESHOLD, 0),
           (rtb200.RT_OPT_BVH_LEAF, 0), (rtb200.RT_OPT_BVH_LEAF, 200), (rtb200.RT_OPT_WF_REFILL, 0), (rtb200.RT_OPT_WF_NODE_MIN, 33),
           (rtb200.RT_OPT_WF_WAVE_MPATHS, -5), (rtb200.RT_OPT_BVH_WIDE, 3), (rtb200.RT_OPT_FLAT_COOP, 9), (rtb200.RT_OPT_POOL_TILES, 99),
           (rtb200.RT_OPT_BVH_SCHED, 2), (99, 0)]
    for opt, v in bad:
        with pytest.raises(rtb200.RtError) as e:
            tracer.set_option(opt, v)
        assert e.value.code == rtb200.RT_ERR_INVALID and "rt_set_option" in str(e.value), (opt, v)
    for opt, v in ((rtb200.RT_OPT_ACCEL, rtb200.RT_ACCEL_AUTO), (rtb200.RT_OPT_BVH_LEAF, 4), (rtb200.RT_OPT_WF_REFILL, 8)):
        tracer.set_option(opt, v)


# ---- library-owned multi-GPU (rt_create_multi / rt_group_*) -----------------------------------------------------------
def _group_against_single(devices, scenes, oracle):
    w, h, spp = 320, 200, 12
    objs = scenes["Scene3_indirect"]
    par = rtb200.default_params(width=w, height=h, mode=rtb200.RT_MODE_PATH, max_bounces=8, seed_lo=11, seed_hi=22)
    g = rtb200.TracerGroup(devices)
    try:
        assert g.size() == len(devices)
        g.set_scene(objs); g.set_camera(rtb200.default_camera()); g.set_params(par)
        g.reset_accumulation()
        g.render_spp(spp // 2); g.render_spp(spp - spp // 2)
        surf = g.resolve_rgba8()
        parts = [m.read_accum() for m in g.members]
        st = g.stats()
        # a second frame after a reset: the group orders "everyone has read my buffer" before the next render touches it
        g.reset_accumulation(); g.render_spp(spp); surf2 = g.resolve_rgba8()
        parts2 = [m.read_accum()[0] for m in g.members]
    finally:
        g.close()
    assert sum(n for _, n in parts) == spp and st.samples == spp and st.paths == w * h * spp
    total = np.zeros((h, w, 4), np.float32)
    for a, _ in parts:
        total = total + a                                 # member order = the fused kernel's summation order
    assert np.array_equal(surf, oracle.resolve_argb8(total, spp))
    total2 = np.zeros((h, w, 4), np.float32)
    for a in parts2:
        total2 = total2 + a
    assert np.array_equal(surf2, oracle.resolve_argb8(total2, spp))
    one = rtb200.PathTracer(devices[0])
    try:
        one.set_scene(objs); one.set_camera(rtb200.default_camera()); one.set_params(par)
        one.reset_accumulation(); one.render_spp(spp)
        ref, _ = one.read_accum()
        seg_one = one.stats().segments
    finally:
        one.close()
    assert np.allclose(ref, total2, rtol=1e-6, atol=1e-6) and st.segments == seg_one      # GPU-count invariant up to summation order


def test_group_of_one_device(scenes, oracle):
    _group_against_single([0], scenes, oracle)


def test_group_across_real_gpus_in_one_process(scenes, oracle):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    _group_against_single(list(range(min(n, 4))), scenes, oracle)


def test_group_rejects_bad_device_lists():
    for devs in ([0, 0], [512]):
        with pytest.raises(rtb200.RtError):
            rtb200.TracerGroup(devs)


def test_resolve_fused_rejects_host_pointers(tracer, scenes):
    """rt_resolve_fused looks its pointers up instead of faulting in the kernel."""
    setup(tracer, scenes["Scene1"], 64, 48)
    tracer.render_spp(1)
    host = np.zeros(64 * 48 * 4, np.float32)
    with pytest.raises(rtb200.RtError) as e:
        tracer.resolve_fused([tracer.accum_device_ptr(), host.ctypes.data], 2, 0, 64 * 48, tracer.argb_device_ptr())
    assert e.value.code == rtb200.RT_ERR_INVALID
    tracer.resolve_fused([tracer.accum_device_ptr()], 1, 0, 64 * 48, tracer.argb_device_ptr())     # still usable afterwards
    tracer.sync()


def test_exchange_across_real_gpus_one_process_per_gpu():
    """rt_exchange_*: device-side flags instead of a collective, one process per GPU (tests/multigpu_fused_check.py)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multigpu_fused_check.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
                          "--master-addr", "127.0.0.1", "--master-port", "29541", script, "--exchange"], capture_output=True, text=True, timeout=600)
    assert "EXCHANGE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


# ---- the C++ host path on the device: host/rt_host.hpp + bin/rt_headless (Raytracer.cpp:364-595 over the C-ABI) ----------
def _read_ppm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P6"
        w, h = [int(x) for x in f.readline().split()]
        assert f.readline().strip() == b"255"
        return np.frombuffer(f.read(), np.uint8).reshape(h, w, 3)


def test_headless_cpp_driver_matches_the_c_abi(tracer, tmp_path, scenes):
    exe = os.path.join(ROOT, "software-raytracer_b200", "bin", "rt_headless")
    if not os.path.exists(exe):
        pytest.skip("bin/rt_headless not built")
    scene_file = str(tmp_path / "Scene1.json")
    assert rtb200.scene_file_write(scene_file, scenes["Scene1"], None, "Scene1") == 0
    w, h, spp = 320, 180, 6
    for preview in (False, True):
        ppm = str(tmp_path / ("p.ppm" if preview else "q.ppm"))
        cmd = [exe, "--scene", scene_file, "--width", str(w), "--height", str(h), "--spp", str(spp), "--bounces", "8", "--scale", "1", "--out", ppm]
        r = subprocess.run(cmd + (["--preview"] if preview else []), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        img = _read_ppm(ppm)
        # the same call sequence through ctypes: reset + the full-resolution overwrite frame, then the remaining samples
        tracer.set_scene(scenes["Scene1"]); tracer.set_camera(rtb200.default_camera())
        tracer.set_params(rtb200.default_params(width=w, height=h, mode=rtb200.RT_MODE_PREVIEW if preview else rtb200.RT_MODE_PATH, max_bounces=8))
        tracer.reset_accumulation()
        tracer.render_spp(1)
        if not preview:
            tracer.render_spp(spp - 1)
        argb = tracer.resolve_rgba8(True)
        want = np.stack([(argb >> 16) & 255, (argb >> 8) & 255, argb & 255], -1).astype(np.uint8)
        assert np.array_equal(img, want), "preview" if preview else "path"
    # the interactive loop (BASELINE configs[4]) runs and reports latencies
    r = subprocess.run([exe, "--scene", scene_file, "--width", "640", "--height", "360", "--interactive", "50", "--scale", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and '"p50_ms"' in r.stdout, r.stdout + r.stderr


def _build_scripted_viewer(tmp_path):
    """host/rt_viewer.cpp (the SDL2 / Dear ImGui front end) linked with the scripted headless double of the
    SDL / ImGui entry points it uses (tests/viewer_stubs/scripted_sdl_imgui.cpp) and the real librt_b200."""
    pkg = os.path.join(ROOT, "software-raytracer_b200")
    exe = str(tmp_path / "rt_viewer_scripted")
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "viewer_stubs"), "-I", os.path.join(pkg, "host"),
                        "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(pkg, "host", "rt_viewer.cpp"),
                        os.path.join(ROOT, "tests", "viewer_stubs", "scripted_sdl_imgui.cpp"), "-L", os.path.join(pkg, "lib"), "-lrt_b200",
                        "-Wl,-rpath," + os.path.join(pkg, "lib")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _run_scripted_viewer(exe, tmp_path, scene_file, w, h, script):
    sp, lp = str(tmp_path / "script.txt"), str(tmp_path / "viewer.log")
    with open(sp, "w") as f:
        f.write(script)
    env = dict(os.environ, RT_VIEWER_SCRIPT=sp, RT_VIEWER_LOG=lp)
    r = subprocess.run([exe, "--scene", scene_file, "--width", str(w), "--height", str(h)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    log = {"title": {}, "name": {}, "surface": {}}
    for line in open(lp):
        _, fr, kind, *rest = line.rstrip("\n").split(" ", 3)
        log[kind][int(fr)] = rest[0] if rest else ""
    frames = {f: int(t.split("ACCUMULATIONFRAMES: ")[1].split(" ")[0]) for f, t in log["title"].items()}
    return log, frames


def test_viewer_main_loop_scripted_on_the_gpu(tracer, tmp_path, scenes):
    """The viewer's REAL main loop (host/rt_viewer.cpp: events, fly camera, inspector, picking, the frame state machine of
    Raytracer.cpp:568-590, rt_render_frame into the streamed surface) on this GPU, with a scripted double in place of SDL2 /
    Dear ImGui (neither is in the image). Static path-mode frames must equal the same call sequence through ctypes bit for bit;
    an interactive script must restart, accumulate, pick, delete, create and save as the reference's loop does."""
    exe = _build_scripted_viewer(tmp_path)
    objs = scenes["Scene1"]
    names = ["object %d" % i for i in range(len(objs))]
    scene_file = str(tmp_path / "Scene1.json")
    assert rtb200.scene_file_write(scene_file, objs, names, "Scene1") == 0
    w, h = 640, 360

    # (1) static camera: switch to path mode in frame 0, six frames, dump the last one
    ppm = str(tmp_path / "static.ppm")
    log, acc = _run_scripted_viewer(exe, tmp_path, scene_file, w, h, "0 button Switch Render Mode\n5 dump %s\n5 quit\n" % ppm)
    assert [acc[f] for f in range(6)] == [1, 2, 3, 4, 5, 6]
    # frame 0 is traced at a quarter of the render scale and overwritten by frame 1 (:576-587); frames 1..5 = 5 samples at render scale 0.5
    tracer.set_scene(objs); tracer.set_camera(rtb200.default_camera())
    tracer.set_params(rtb200.default_params(width=w, height=h, mode=rtb200.RT_MODE_PATH, max_bounces=2))     # MAXBOUNCES = 2 (Raytracer.cpp:33)
    tracer.reset_accumulation()
    tracer.set_pixel_step(rtb200.reference_pixel_step(0.5, 1.0), rtb200.reference_strip_columns(w))
    for _ in range(5):
        tracer.render_spp(1)
    argb = tracer.resolve_rgba8(True)
    tracer.set_pixel_step(1, 0)
    want = np.stack([(argb >> 16) & 255, (argb >> 8) & 255, argb & 255], -1).astype(np.uint8)
    assert np.array_equal(_read_ppm(ppm), want)
    assert len(set(log["surface"][f] for f in range(6))) == 6            # every frame shows a different image (noise averages out)

    # (2) interaction
    tracer.set_params(rtb200.default_params(width=w, height=h, mode=rtb200.RT_MODE_PREVIEW, max_bounces=2))
    centre = tracer.pick(w // 2, h // 2)
    assert centre >= 0
    saved = str(tmp_path / "saved.json")
    script = "\n".join([
        "3 button Switch Render Mode",                       # preview -> path mode: restart
        "8 click %d %d" % (w // 2, h // 2),                  # select what the centre pixel sees
        "12 rmb down", "13 motion 50 0", "14 rmb up",        # look around: restart while the button is held
        "18 key DELETE",                                     # remove the selected object: restart
        "21 menu Sphere", "22 setfloat Sphere Radius=0.75",  # create a sphere 5 units ahead, edit it in the inspector
        "24 hold W on", "25 hold W off",                     # fly forward for one frame
        "27 key P", "29 key P",                              # pause / resume
        "30 settext Save path=%s" % saved, "30 menu Save",
        "31 quit"]) + "\n"
    log, acc = _run_scripted_viewer(exe, tmp_path, scene_file, w, h, script)
    assert [acc[f] for f in range(0, 3)] == [1, 1, 1]                    # preview mode never accumulates (:589)
    assert [acc[f] for f in range(3, 12)] == list(range(1, 10))          # path mode: restart, then one more frame each
    assert log["name"].get(8) is None and all(log["name"][f] == names[centre] for f in range(9, 18))   # the inspector shows the picked object from the next frame on
    assert acc[12] == 1 and acc[13] == 1 and acc[14] == 2                # right button held: every frame restarts (:390-394)
    assert log["surface"][11] != log["surface"][14]                      # the camera turned
    assert acc[18] == 1 and 19 not in log["name"]                        # delete: scene re-submitted, nothing selected
    assert acc[21] == 1 and log["name"][21] == "" and log["name"][22] == ""     # create: selected at once, no name yet (as in the reference)
    assert acc[22] == 1                                                  # the radius edit re-submits the scene
    assert acc[24] == 1 and acc[25] == 2                                 # camera moved for one frame
    assert acc[27] == acc[26] and acc[28] == acc[26] and acc[29] == acc[26] + 1    # paused frames render nothing
    rc, objs2, err = rtb200.scene_file_read(saved)
    assert rc == 0, err
    assert len(objs2) == len(objs) and objs2[-1]["radius"] == np.float32(0.75)
    kept = [i for i in range(len(objs)) if i != centre]
    assert np.array_equal(objs2[:-1], np.ascontiguousarray(objs, rtb200.OBJECT_DTYPE)[kept])


def test_headless_cpp_driver_on_several_gpus(tmp_path, scenes):
    """`rt_headless --gpus N`: the library-owned group from a C++ host - no torch, no IPC (Raytracer.cpp:331-342, 598-607)."""
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    exe = os.path.join(ROOT, "software-raytracer_b200", "bin", "rt_headless")
    scene_file = str(tmp_path / "Scene1.json")
    assert rtb200.scene_file_write(scene_file, scenes["Scene1"], None, "Scene1") == 0
    w, h, spp = 320, 180, 16
    ppm = str(tmp_path / "g.ppm")
    r = subprocess.run([exe, "--scene", scene_file, "--width", str(w), "--height", str(h), "--spp", str(spp), "--bounces", "8", "--gpus", str(n), "--out", ppm],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and '"gpus": %d' % n in r.stdout, r.stdout + r.stderr
    g = rtb200.TracerGroup(list(range(n)))
    try:
        g.set_scene(scenes["Scene1"]); g.set_camera(rtb200.default_camera())
        g.set_params(rtb200.default_params(width=w, height=h, mode=rtb200.RT_MODE_PATH, max_bounces=8))
        g.reset_accumulation(); g.render_spp(spp)
        argb = g.resolve_rgba8(True)
    finally:
        g.close()
    want = np.stack([(argb >> 16) & 255, (argb >> 8) & 255, argb & 255], -1).astype(np.uint8)
    assert np.array_equal(_read_ppm(ppm), want)


@pytest.mark.parametrize("reuse", [1, 0])
def test_streaming_pipeline_is_bit_identical_on_small_bvh_scenes(tracer, scenes, reuse):
    """RT_PIPELINE_STREAM over a BVH staged in shared memory (bundled scenes, sphere + cube mix) and read from global memory
    (3000 primitives), with and without primary-hit reuse, max_bounces 0 and 8: the megakernel's bits and counts."""
    cases = [(scenes["Scene1"], 333, 77, 8), (scenes["Scene3_indirect"], 160, 120, 8), (scenes["Scene_indirect"], 160, 120, 0),
             (synthetic_spheres(3000, cubes_every=9), 256, 144, 8)]
    try:
        tracer.set_option(rtb200.RT_OPT_ACCEL, rtb200.RT_ACCEL_BVH)
        tracer.set_option(rtb200.RT_OPT_PRIMARY_REUSE, reuse)
        for i, (objs, w, h, mb) in enumerate(cases):
            res = {}
            for pipe in (rtb200.RT_PIPELINE_REGEN, rtb200.RT_PIPELINE_STREAM):
                tracer.set_option(rtb200.RT_OPT_PIPELINE, pipe)
                setup(tracer, objs, w, h, config3_camera(rtb200.default_camera) if i == 3 else None, max_bounces=mb)
                tracer.render_spp(5); tracer.render_spp(2)
                st = tracer.stats()
                assert st.pipeline == pipe
                res[pipe] = (tracer.read_accum()[0], st.segments, st.traced_segments, st.paths)
            a, b = res[rtb200.RT_PIPELINE_REGEN], res[rtb200.RT_PIPELINE_STREAM]
            assert np.array_equal(bits(a[0]), bits(b[0])) and a[1:] == b[1:], (i, reuse)
    finally:
        tracer.set_option(rtb200.RT_OPT_PIPELINE, rtb200.RT_PIPELINE_AUTO)
        tracer.set_option(rtb200.RT_OPT_ACCEL, rtb200.RT_ACCEL_AUTO)
        tracer.set_option(rtb200.RT_OPT_PRIMARY_REUSE, 1)


# ---- rt_render_frame: the frame resolved (and streamed to the host surface) by the render kernel itself ----------------------------
@pytest.mark.parametrize("scene,w,h", [("Scene1", 333, 77), ("Scene3_indirect", 640, 360), ("Scene1", 1283, 721)])
def test_render_frame_equals_render_then_resolve(tracer, scenes, scene, w, h):
    """One call per progressive frame (Raytracer.cpp:223-257: renderArea resolves every pixel as it traces it) == rt_render_spp +
    rt_resolve_rgba8, bit for bit: surface AND accumulation buffer, frame after frame, page-locked and pageable surfaces, both row
    orders, widths that cut tiles (partial 8x4 tiles and a last chunk of one tile), 1 spp (fused) and 4 spp (two steps inside)."""
    pinned, _owner = rtb200.host_surface(w, h)
    for flip in (True, False):
        setup(tracer, scenes[scene], w, h)
        want = []
        for n in (1, 1, 1, 4, 1, 2):
            tracer.render_spp(n)
            want.append(tracer.resolve_rgba8(flip).copy())
        acc_want, n_want = tracer.read_accum()
        setup(tracer, scenes[scene], w, h)
        for i, n in enumerate((1, 1, 1, 4, 1, 2)):
            pinned[:] = 0xdeadbeef
            got = tracer.render_frame(n, flip, pinned)
            assert np.array_equal(got, want[i]), (flip, i, int((got != want[i]).sum()))
        acc, n_got = tracer.read_accum()
        assert n_got == n_want and np.array_equal(bits(acc), bits(acc_want))
        setup(tracer, scenes[scene], w, h)               # pageable destination: fused resolve + one copy
        for i, n in enumerate((1, 1, 1)):
            got = tracer.render_frame(n, flip)
            assert np.array_equal(got, want[i]), (flip, i)


def test_render_frame_sky_only_and_other_modes(tracer, scenes):
    """Every pixel a miss (finished the moment it is picked up), preview mode and block-filled frames (not fused: same result)."""
    w, h = 200, 50
    pinned, _owner = rtb200.host_surface(w, h)
    setup(tracer, scenes["Scene1"][:0], w, h)
    tracer.render_spp(1)
    want = tracer.resolve_rgba8().copy()
    setup(tracer, scenes["Scene1"][:0], w, h)
    assert np.array_equal(tracer.render_frame(1, True, pinned), want)
    setup(tracer, scenes["Scene1"], w, h)
    tracer.set_pixel_step(4, 16)
    try:
        tracer.render_spp(1)
        want = tracer.resolve_rgba8().copy()
        tracer.reset_accumulation()
        assert np.array_equal(tracer.render_frame(1, True, pinned), want)
    finally:
        tracer.set_pixel_step(1, 0)
