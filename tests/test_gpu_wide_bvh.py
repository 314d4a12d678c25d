"""GPU (B200): the 8-wide BVH with quantised child boxes (csrc/bvh_wide.h) gives the same bits as the binary BVH and
as the brute-force object loop - primary AOVs, arbitrary rays, accumulated radiance, segment counts - in both
pipelines. The wide form only produces candidates; hits are decided by the strict reference intersectors."""
import numpy as np
import pytest

import rtb200
from conftest import SCENES, make_camera
from rtb200.scenes import heightfield_mesh, mesh_scene, synthetic_spheres
from test_gpu_parity import bits, setup, _mesh_cam

pytestmark = pytest.mark.gpu


def _reset(tracer):
    tracer.set_option(rtb200.RT_OPT_BVH_WIDE, 0)
    tracer.set_option(rtb200.RT_OPT_ACCEL, rtb200.RT_ACCEL_AUTO)
    tracer.set_option(rtb200.RT_OPT_PIPELINE, rtb200.RT_PIPELINE_AUTO)
    tracer.set_option(rtb200.RT_OPT_BVH_SCHED, 0)


@pytest.mark.parametrize("scene", SCENES)
def test_wide_bvh_on_bundled_scenes(tracer, scenes, meta, scene):
    from conftest import sha
    try:
        tracer.set_option(rtb200.RT_OPT_ACCEL, rtb200.RT_ACCEL_BVH)
        tracer.set_option(rtb200.RT_OPT_BVH_WIDE, 2)                 # force the wide form on a small scene
        for cam_name in ("default", "rotated"):
            m = meta["aov"]["%s_640x480_%s" % (scene, cam_name)]
            setup(tracer, scenes[scene], 640, 480, make_camera(rtb200.RtCamera, meta, cam_name == "rotated"))
            ids, t, nrm, pt = tracer.read_aov()
            assert sha(ids) == m["ids_sha256"] and sha(t) == m["t_sha256"]
            assert sha(nrm) == m["normal_sha256"] and sha(pt) == m["point_sha256"]
        out = {}
        for wide, accel, pipe in ((2, rtb200.RT_ACCEL_BVH, rtb200.RT_PIPELINE_REGEN), (2, rtb200.RT_ACCEL_BVH, rtb200.RT_PIPELINE_WAVEFRONT),
                                  (0, rtb200.RT_ACCEL_BVH, rtb200.RT_PIPELINE_REGEN), (0, rtb200.RT_ACCEL_BRUTE, rtb200.RT_PIPELINE_REGEN)):
            tracer.set_option(rtb200.RT_OPT_BVH_WIDE, wide); tracer.set_option(rtb200.RT_OPT_ACCEL, accel)
            tracer.set_option(rtb200.RT_OPT_PIPELINE, pipe)
            setup(tracer, scenes[scene], 160, 120)
            tracer.render_spp(16)
            st = tracer.stats()
            assert st.accel == accel and st.pipeline == pipe
            out[(wide, accel, pipe)] = (tracer.read_accum()[0], st.segments)
        ref = out[(0, rtb200.RT_ACCEL_BRUTE, rtb200.RT_PIPELINE_REGEN)]
        for k, v in out.items():
            assert np.array_equal(bits(v[0]), bits(ref[0])) and v[1] == ref[1], k
    finally:
        _reset(tracer)


def _rays(n, seed, lo, hi):
    rng = np.random.default_rng(seed)
    org = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[::17, 0] = 0.0; d[5::29, 1] = 0.0; d[7::31, 2] = -0.0           # zero components: the clamped reciprocal path
    d[11::97] = (0.0, 0.0, 1.0); d[13::101] = (0.0, -1.0, 0.0)        # axis-parallel rays
    d[np.linalg.norm(d, axis=1) == 0] = (1.0, 0.0, 0.0)               # (a row hit by all three zeroing patterns)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    # rt_trace_rays sends a batch with ANY non-unit direction through the brute-force loop: make sure this one qualifies
    assert np.all(np.abs((d.astype(np.float32) ** 2).sum(axis=1) - 1.0) <= 1e-6)
    return org, d


def test_wide_bvh_large_sphere_cube_scene(tracer):
    objs = synthetic_spheres(6000, cubes_every=7)
    cam = rtb200.default_camera(60)
    cam.pos[0], cam.pos[1], cam.pos[2] = 0.0, 8.0, -20.0
    org, d = _rays(60000, 3, (-40, 0.05, -40), (40, 12, 40))
    org[::5] = objs["pos"][np.arange(0, len(org[::5])) % len(objs)]   # origins at object centres: inside spheres and cubes
    res = {}
    try:
        for key, wide, accel, pipe in (("wide", 1, rtb200.RT_ACCEL_BVH, rtb200.RT_PIPELINE_REGEN), ("wide_wf", 1, rtb200.RT_ACCEL_BVH, rtb200.RT_PIPELINE_WAVEFRONT),
                                       ("wide_wf_batch", 1, rtb200.RT_ACCEL_BVH, rtb200.RT_PIPELINE_WAVEFRONT),
                                       ("bvh2", 0, rtb200.RT_ACCEL_BVH, rtb200.RT_PIPELINE_REGEN), ("brute", 0, rtb200.RT_ACCEL_BRUTE, rtb200.RT_PIPELINE_REGEN)):
            tracer.set_option(rtb200.RT_OPT_BVH_WIDE, wide); tracer.set_option(rtb200.RT_OPT_ACCEL, accel)
            tracer.set_option(rtb200.RT_OPT_PIPELINE, pipe)
            tracer.set_option(rtb200.RT_OPT_BVH_SCHED, 1 if key == "wide_wf_batch" else 0)
            setup(tracer, objs, 256, 144, cam)
            aov = tracer.read_aov()
            rays = tracer.trace_rays(org, d)
            tracer.render_spp(5)
            st = tracer.stats()
            assert st.accel == accel and st.pipeline == pipe
            res[key] = (aov, rays, tracer.read_accum()[0], st.segments)
    finally:
        _reset(tracer)
    ref = res["brute"]
    assert (ref[0][0] >= 0).mean() > 0.5 and (ref[1][0] >= 0).mean() > 0.3
    for k, v in res.items():
        for x, y in zip(v[0] + v[1], ref[0] + ref[1]):
            hit = np.asarray(ref[0][0] if x.shape[:2] == ref[0][0].shape else ref[1][0]) >= 0
            assert np.array_equal(bits(x)[hit], bits(y)[hit]), k
        assert np.array_equal(v[0][0], ref[0][0]) and np.array_equal(v[1][0], ref[1][0]), k
        assert np.array_equal(bits(v[2]), bits(ref[2])) and v[3] == ref[3], k


def test_wide_bvh_mesh(tracer):
    v, tr = heightfield_mesh(96, 64, seed=7)                           # 12 160 triangles
    objs = mesh_scene()
    org, d = _rays(40000, 9, (-10, 0.0, -6), (10, 4, 6))
    res = {}
    try:
        for key, wide, pipe in (("wide", 1, rtb200.RT_PIPELINE_REGEN), ("wide_wf", 1, rtb200.RT_PIPELINE_WAVEFRONT), ("bvh2", 0, rtb200.RT_PIPELINE_REGEN)):
            tracer.set_option(rtb200.RT_OPT_BVH_WIDE, wide); tracer.set_option(rtb200.RT_OPT_PIPELINE, pipe)
            setup(tracer, objs, 200, 120, _mesh_cam(rtb200.RtCamera), max_bounces=5)
            tracer.set_mesh(0, v, tr)
            aov = tracer.read_aov()
            rays = tracer.trace_rays(org, d)
            tracer.render_spp(4)
            res[key] = (aov, rays, tracer.read_accum()[0], tracer.stats().segments)
    finally:
        _reset(tracer)
        tracer.set_scene(mesh_scene()[1:])
    ref = res["bvh2"]                                                  # itself pinned to in-order brute force + oracle (test_gpu_parity)
    assert (ref[0][0] == 0).mean() > 0.2
    for k, val in res.items():
        assert np.array_equal(val[0][0], ref[0][0]) and np.array_equal(val[1][0], ref[1][0]), k
        for x, y, ids in [(a, b, ref[0][0]) for a, b in zip(val[0][1:], ref[0][1:])] + [(a, b, ref[1][0]) for a, b in zip(val[1][1:], ref[1][1:])]:
            assert np.array_equal(bits(x)[ids >= 0], bits(y)[ids >= 0]), k
        assert np.array_equal(bits(val[2]), bits(ref[2])) and val[3] == ref[3], k


def test_wavefront_wave_size_does_not_change_the_sum(tracer):
    """Samples per wave (RT_OPT_WF_WAVE_MPATHS) only change how the work is batched: per pixel the samples are still added in
    sample order, so tiny waves, the default and the megakernel give the same bits; RT_PIPELINE_AUTO (which may pick either
    pipeline for a long call, and always the megakernel for short ones) as well."""
    objs = synthetic_spheres(3000, cubes_every=9)
    cam = rtb200.default_camera(60)
    cam.pos[0], cam.pos[1], cam.pos[2] = 0.0, 8.0, -20.0
    out = {}
    try:
        for key, pipe, wave in (("regen", rtb200.RT_PIPELINE_REGEN, 0), ("wf_default", rtb200.RT_PIPELINE_WAVEFRONT, 0),
                                ("wf_tiny", rtb200.RT_PIPELINE_WAVEFRONT, 1), ("auto", rtb200.RT_PIPELINE_AUTO, 0)):
            tracer.set_option(rtb200.RT_OPT_PIPELINE, pipe); tracer.set_option(rtb200.RT_OPT_WF_WAVE_MPATHS, wave)
            setup(tracer, objs, 512, 288, cam)                       # 147 456 pixels: 1 M paths = 7 samples per wave
            tracer.render_spp(3); tracer.render_spp(37)
            st = tracer.stats()
            if key == "regen" or key.startswith("wf"):
                assert st.pipeline == pipe
            out[key] = (tracer.read_accum()[0], st.segments, st.paths)
    finally:
        tracer.set_option(rtb200.RT_OPT_WF_WAVE_MPATHS, 0)
        _reset(tracer)
    for k, v in out.items():
        assert np.array_equal(bits(v[0]), bits(out["regen"][0])) and v[1:] == out["regen"][1:], k
