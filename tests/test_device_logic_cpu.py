"""CPU: the product's DEVICE functions (csrc/rt_device.cuh: raygen, strict intersectors, BVH candidate
traversal, shading, Philox) and host packing / BVH builder, compiled for the host through
tests/host_emu/cuda_shim.h, against the oracle and the reference-generated goldens. This checks the
device-side logic where there is no GPU; the -m gpu tests check the code nvcc actually generates."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import rtb200
from conftest import ROOT, SCENES, SEED, make_camera, sha
from oracle_py import OrcCamera

EMU_DIR = os.path.join(ROOT, "tests", "host_emu")
EMU_SO = os.path.join(EMU_DIR, "libemu.so")


def build_emu():
    """g++ build of the device code for the host (tests/host_emu/emu.cpp + the host builders); returns the library path."""
    srcs = [os.path.join(EMU_DIR, "emu.cpp"), os.path.join(ROOT, "software-raytracer_b200", "csrc", "bvh_build.cpp"),
            os.path.join(ROOT, "software-raytracer_b200", "csrc", "bvh_wide.cpp"),
            os.path.join(ROOT, "software-raytracer_b200", "csrc", "flat_build.cpp"),
            os.path.join(ROOT, "software-raytracer_b200", "csrc", "mesh.cpp")]
    hdrs = [os.path.join(ROOT, "software-raytracer_b200", "csrc", f) for f in ("rt_device.cuh", "rt_bvh_lane.cuh")]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(f) > os.path.getmtime(EMU_SO) for f in srcs + hdrs):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I" + os.path.join(ROOT, "include"),
                               "-o", EMU_SO] + srcs)
    return EMU_SO


@pytest.fixture(scope="session")
def emu():
    lib = C.CDLL(build_emu())
    lib.emu_render.restype = C.c_longlong
    lib.emu_render_mesh.restype = C.c_longlong

    def render(objs, cam, par, accel, s0, n, aov=False, mesh=None):
        w, h = par.width, par.height
        objs = np.ascontiguousarray(objs, rtb200.OBJECT_DTYPE)
        out = np.zeros((h, w, 3), np.float32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        ids = np.zeros((h, w), np.int32); t = np.zeros((h, w), np.float32)
        nrm = np.zeros((h, w, 3), np.float32); pt = np.zeros((h, w, 3), np.float32)
        if mesh is None:
            segs = lib.emu_render(p(objs), len(objs), C.byref(cam), C.byref(par), accel, C.c_uint32(s0), n, p(out),
                                  p(ids) if aov else None, p(t), p(nrm), p(pt))
        else:
            oi, v, tr = mesh
            v = np.ascontiguousarray(v, np.float32).reshape(-1, 3); tr = np.ascontiguousarray(tr, np.int32).reshape(-1, 3)
            segs = lib.emu_render_mesh(p(objs), len(objs), C.byref(cam), C.byref(par), accel, C.c_uint32(s0), n, p(out),
                                       p(ids) if aov else None, p(t), p(nrm), p(pt), p(v), len(v), p(tr), len(tr), oi)
        return out, segs, (ids, t, nrm, pt)
    return render


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("accel", [0, 1, 2, 3])
def test_device_logic_matches_reference_goldens(emu, scenes, golden, meta, scene, accel):
    z = golden("radiance_philox")
    for cam_name, mb in (("default", 8), ("default", 2), ("rotated", 8)):
        cam = make_camera(rtb200.RtCamera, meta, cam_name == "rotated")
        par = rtb200.default_params(width=64, height=48, mode=0, max_bounces=mb, seed_lo=SEED[0], seed_hi=SEED[1])
        out, segs, _ = emu(scenes[scene], cam, par, accel, 0, 4)
        want = z["%s_%s_mb%d_sum" % (scene, cam_name, mb)]
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), (cam_name, mb)   # glibc powf on both sides: bit-equal
    for cam_name in ("default", "rotated"):
        m = meta["aov"]["%s_160x120_%s" % (scene, cam_name)]
        cam = make_camera(rtb200.RtCamera, meta, cam_name == "rotated")
        par = rtb200.default_params(width=160, height=120, mode=0, max_bounces=8)
        _, _, (ids, t, nrm, pt) = emu(scenes[scene], cam, par, accel, 0, 0, aov=True)
        assert sha(ids) == m["ids_sha256"] and sha(t) == m["t_sha256"] and sha(nrm) == m["normal_sha256"] and sha(pt) == m["point_sha256"]


@pytest.mark.parametrize("accel", [1, 3])
def test_device_logic_bvh_equals_brute_on_random_scene(emu, accel):
    rng = np.random.default_rng(7)
    n = 600
    o = np.zeros(n + 1, rtb200.OBJECT_DTYPE)
    o["type"] = 1
    o["pos"][:n] = rng.uniform([-12, 0, 4], [12, 8, 40], (n, 3)).astype(np.float32)
    o["radius"][:n] = rng.uniform(0.1, 0.8, n).astype(np.float32)
    o["base"] = rng.uniform(0.1, 0.9, (n + 1, 3)).astype(np.float32); o["spec_color"] = 1
    o["spec_amount"][::3] = 1; o["smoothness"][::3] = 0.9
    o["type"][5::11] = 2; o["half"][5::11] = rng.uniform(0.2, 0.7, (len(o["half"][5::11]), 3)).astype(np.float32)
    o["emissive"][::40] = 20
    o["pos"][n] = [0, -500, 20]; o["radius"][n] = 500
    cam = rtb200.default_camera(60); cam.pos[1] = 3.0; cam.pos[2] = -6.0
    par = rtb200.default_params(width=96, height=64, mode=0, max_bounces=6, seed_lo=5, seed_hi=6)
    a, sa, aa = emu(o, cam, par, 0, 3, 3, aov=True)
    b, sb, ab = emu(o, cam, par, accel, 3, 3, aov=True)      # 1: binary BVH, 3: 8-wide quantised BVH
    assert sa == sb and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for x, y in zip(aa, ab):
        assert np.array_equal(np.ascontiguousarray(x).view(np.uint32), np.ascontiguousarray(y).view(np.uint32))
    assert (aa[0] >= 0).mean() > 0.5


@pytest.mark.parametrize("n,spread,seed", [(250, 12.0, 1), (120, 2.0, 2), (40, 0.6, 3), (9, 5.0, 4), (3, 1.0, 5)])
def test_device_logic_flat_equals_brute_on_random_scenes(emu, n, spread, seed):
    """flat two-level accelerator (conservative culls + strict tests) vs the brute-force loop: ids, t, n, p,
    radiance and segment counts bit-equal. Small `spread` packs overlapping spheres so lanes overflow the
    8-entry candidate queue (exact fallback) and origins sit inside several spheres (negative t)."""
    rng = np.random.default_rng(seed)
    o = np.zeros(n + 1, rtb200.OBJECT_DTYPE)
    o["type"] = 1
    o["pos"][:n] = rng.uniform([-spread, 0, 4], [spread, spread, 4 + 2 * spread], (n, 3)).astype(np.float32)
    o["radius"][:n] = rng.uniform(0.1, 0.8, n).astype(np.float32)
    o["base"] = rng.uniform(0.1, 0.9, (n + 1, 3)).astype(np.float32); o["spec_color"] = 1
    o["spec_amount"][::3] = 1; o["smoothness"][::3] = 0.9
    cubes = np.arange(2, n, 7)
    o["type"][cubes] = 2; o["half"][cubes] = rng.uniform(0.2, 0.7, (len(cubes), 3)).astype(np.float32)
    o["emissive"][::10] = 20
    o["pos"][n] = [0, -500, 20]; o["radius"][n] = 500
    cam = rtb200.default_camera(60); cam.pos[1] = 0.5 * spread; cam.pos[2] = -2.0
    par = rtb200.default_params(width=96, height=64, mode=0, max_bounces=6, seed_lo=5, seed_hi=6)
    a, sa, aa = emu(o, cam, par, 0, 3, 3, aov=True)
    b, sb, ab = emu(o, cam, par, 2, 3, 3, aov=True)
    assert sb >= 0, "flat accelerator not usable for this scene"
    assert sa == sb and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for x, y in zip(aa, ab):
        assert np.array_equal(np.ascontiguousarray(x).view(np.uint32), np.ascontiguousarray(y).view(np.uint32))


def test_device_logic_mesh_bvh_equals_brute_and_oracle(emu, oracle):
    """mesh extension: triangles through the BVH vs the in-order brute-force loop vs the oracle's restatement
    (ids, t, normals, points bit-equal; radiance bit-equal between the two back ends)."""
    from rtb200.scenes import heightfield_mesh, mesh_scene
    v, tr = heightfield_mesh(24, 16, seed=3)
    objs = mesh_scene()
    cam = rtb200.default_camera(55); cam.pos[1] = 1.5; cam.pos[2] = -1.0
    par = rtb200.default_params(width=96, height=64, mode=0, max_bounces=4, seed_lo=5, seed_hi=6)
    a, sa, aa = emu(objs, cam, par, 0, 0, 3, aov=True, mesh=(0, v, tr))
    for accel in (1, 3):                                     # binary BVH, 8-wide quantised BVH
        b, sb, ab = emu(objs, cam, par, accel, 0, 3, aov=True, mesh=(0, v, tr))
        assert sa == sb and np.array_equal(a.view(np.uint32), b.view(np.uint32)), accel
        for x, y in zip(aa, ab):
            assert np.array_equal(np.ascontiguousarray(x).view(np.uint32), np.ascontiguousarray(y).view(np.uint32)), accel
    assert (aa[0] == 0).mean() > 0.2                         # the mesh is visible
    ocam = OrcCamera(); ocam.right[0] = 1; ocam.up[1] = 1; ocam.forward[2] = 1; ocam.fov_deg = 55; ocam.pos[1] = 1.5; ocam.pos[2] = -1.0
    oracle.set_triangles(objs, {0: (v, tr)})
    try:
        oid, ot, on, op = oracle.primary_aov(objs, ocam, 96, 64)
        op_ = oracle.default_params(width=96, height=64, max_bounces=4, mode=0, seed_lo=5, seed_hi=6)
        want, _, segs = oracle.render(objs, ocam, op_, 0, 3)
    finally:
        oracle.set_triangles(objs, {})
    assert np.array_equal(aa[0], oid)
    hit = oid >= 0
    assert np.array_equal(aa[1][hit].view(np.uint32), ot[hit].view(np.uint32))
    assert np.array_equal(aa[2][hit].view(np.uint32), on[hit].view(np.uint32))
    assert np.array_equal(aa[3][hit].view(np.uint32), op[hit].view(np.uint32))
    assert segs == sa and np.array_equal(a.view(np.uint32), want.view(np.uint32))


def test_wide_bvh_makes_about_half_the_node_visits(emu):
    """DESIGN.md's claim about the 8-wide quantised BVH, from the emulated traversal loops' own counters: roughly 2.2x fewer
    node visits per ray than the binary tree on a few thousand random spheres (each visit costs ~4x the instructions, which
    is why it is not the default)."""
    lib = C.CDLL(EMU_SO)
    rng = np.random.default_rng(5)
    n = 4000
    o = np.zeros(n + 1, rtb200.OBJECT_DTYPE); o["type"] = 1
    o["pos"] = rng.uniform(-20, 20, (n + 1, 3)).astype(np.float32); o["pos"][:, 1] = rng.uniform(0.2, 6, n + 1)
    o["radius"] = rng.uniform(0.1, 0.6, n + 1).astype(np.float32)
    o["pos"][n] = (0, -1000, 0); o["radius"][n] = 1000
    o["base"] = 0.7; o["spec_color"] = 1
    cam = rtb200.default_camera(60); cam.pos[1] = 3.0; cam.pos[2] = -26.0
    par = rtb200.default_params(width=64, height=48, mode=0, max_bounces=4, seed_lo=5, seed_hi=6)
    stats = (C.c_longlong * 6)()
    lib.emu_bvh_stats(stats)                                  # reset
    a, sa, _ = emu(o, cam, par, 1, 0, 2)
    lib.emu_bvh_stats(stats); b2_visits, b2_prims = stats[2], stats[5]
    b, sb, _ = emu(o, cam, par, 3, 0, 2)
    lib.emu_bvh_stats(stats); w_visits, w_prims = stats[0], stats[1]
    assert sa == sb and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert b2_visits > 0 and 0.3 * b2_visits < w_visits < 0.6 * b2_visits, (b2_visits, w_visits)
    assert w_prims < 2.0 * b2_prims, (b2_prims, w_prims)


@pytest.mark.parametrize("use_slots", [0, 1, 2])
@pytest.mark.parametrize("leaf_bias", [0, 2, 8])
def test_lane_state_machine_equals_the_per_ray_loop_under_random_schedules(emu, leaf_bias, use_slots):
    """csrc/rt_bvh_lane.cuh (the traversal state machine of k_wf_intersect_bvh and k_wf_stream with its sentinel stack), compiled
    for the host: whatever the interleaving of node steps and postponed leaf steps, the hit is closest_hit_bvh()'s bit for bit
    and the stack is back at its sentinel. Spheres + cubes + a triangle mesh, origins inside and outside the scene; with the leaf
    primitives read through refs[] (0), through the leaf-ordered 64-byte slots (1, bvh_build.h build_leaf_slots), and with the
    32-byte nodes whose child planes are quantised to 16 bits (2, HostQNodes): the decoded boxes must contain the float ones for
    every ray, or a hit would be lost."""
    from rtb200.scenes import synthetic_spheres, heightfield_mesh, mesh_scene
    lib = C.CDLL(EMU_SO)
    lib.emu_lane_schedules.restype = C.c_int
    rng = np.random.default_rng(31 + leaf_bias)
    n = 1500
    cases = []
    o = synthetic_spheres(700, seed=9, cubes_every=7)
    org = rng.uniform([-60, -2, -10], [60, 25, 115], (n, 3)).astype(np.float32)
    cases.append((o, None, org))
    v, tr = heightfield_mesh(40, 24, seed=6)
    cases.append((mesh_scene(), (v, tr), rng.uniform([-9, -2.5, -1], [9, 5, 12], (n, 3)).astype(np.float32)))
    for objs, mesh, org in cases:
        d = rng.normal(size=(n, 3))
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        d[::50, 0] = 0; d[::50] /= np.linalg.norm(d[::50], axis=1, keepdims=True)       # exactly-zero direction components
        d = d.astype(np.float32)
        objs = np.ascontiguousarray(objs, rtb200.OBJECT_DTYPE)
        steps = (C.c_longlong * 2)()
        p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        mv = np.ascontiguousarray(mesh[0], np.float32) if mesh else None
        mt = np.ascontiguousarray(mesh[1], np.int32) if mesh else None
        bad = lib.emu_lane_schedules(p(objs), len(objs), p(mv), len(mv) if mesh else 0, p(mt), len(mt) if mesh else 0, 0 if mesh else -1,
                                     p(org), p(d), n, 77 + leaf_bias, leaf_bias, use_slots, steps)
        assert bad == 0, (leaf_bias, use_slots, bad)
        assert steps[0] > 2 * n and steps[1] > n / 4


def test_two_word_minimum_protocol_of_the_pooled_traversal():
    """closest_hit_flat_coop (csrc/rt_trace.cuh) returns a warp's strict test results to the owning lanes as the minimum of (distance
    bits, object id) kept in TWO 32-bit shared-memory words per owner: per pass, atomicMin on the distance word; a candidate that LOWERED
    it resets the id word; candidates that hold the owner's smallest distance atomicMin their id. A model of that protocol with the
    atomics applied in random order must give the lexicographic minimum for any stream of passes (ties on distance included)."""
    rng = np.random.default_rng(7)
    for trial in range(300):
        n_owner = int(rng.integers(1, 6))
        key_t = [0xffffffff] * n_owner
        key_id = [0xffffffff] * n_owner
        seen = [[] for _ in range(n_owner)]
        for _ in range(int(rng.integers(1, 6))):                      # passes, separated by __syncwarp
            lanes = []
            for _ in range(int(rng.integers(0, 33))):
                lanes.append((int(rng.integers(0, n_owner)), int(rng.integers(0, 4)) * 1000 + 5, int(rng.integers(0, 200))))   # few distinct distances: many ties
            for ow, ob, idc in lanes:
                seen[ow].append((ob, idc))
            before = {}
            for k in rng.permutation(len(lanes)):                     # phase 1: read, then atomicMin - interleaved in any order
                ow, ob, idc = lanes[k]
                before[k] = key_t[ow]
                if ob < before[k]:
                    key_t[ow] = min(key_t[ow], ob)
            now = {k: key_t[lanes[k][0]] for k in range(len(lanes))}   # after the first __syncwarp
            for k in rng.permutation(len(lanes)):                     # a new smallest distance: its ids start afresh
                ow, ob, idc = lanes[k]
                if now[k] < before[k] and ob == now[k]:
                    key_id[ow] = 0xffffffff
            for k in rng.permutation(len(lanes)):                     # after the second __syncwarp
                ow, ob, idc = lanes[k]
                if ob == now[k]:
                    key_id[ow] = min(key_id[ow], idc)
        for ow in range(n_owner):
            want = min(seen[ow]) if seen[ow] else (0xffffffff, 0xffffffff)
            assert (key_t[ow], key_id[ow]) == want, (trial, ow)
