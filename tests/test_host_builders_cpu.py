"""CPU: the product's HOST builders (flat accelerator, triangle records, OBJ subset), compiled for the host with the
emulation library. Invariants the device code relies on, checked without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

import rtb200
from conftest import SCENES
from test_device_logic_cpu import emu, EMU_SO  # noqa: F401  (session fixture builds the library)


@pytest.fixture(scope="module")
def lib(emu):  # noqa: F811
    return C.CDLL(EMU_SO)


def flat_info(lib, objs, extent=0.0, offset=1e-5):
    objs = np.ascontiguousarray(objs, rtb200.OBJECT_DTYPE)
    counts = (C.c_int * 4)()
    cull = np.zeros((4096, 4), np.float32); slots = np.zeros(4096, np.uint8); boxes = np.zeros((256, 8), np.float32)
    ki = (C.c_float * 2)()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ok = lib.emu_flat_info(p(objs), len(objs), C.c_float(extent), C.c_float(offset), counts, p(cull), 4096, p(slots), 4096, p(boxes), 256, ki)
    nc, ncube, ns, nrec = list(counts)
    return ok, nc, ncube, ns, cull[:nrec], slots[:8 * nc + ns], boxes[:nc + ncube], ki[0], ki[1]


@pytest.mark.parametrize("scene", SCENES)
def test_flat_builder_invariants_on_bundled_scenes(lib, scenes, scene):
    objs = scenes[scene]
    ok, nc, ncube, ns, cull, slots, boxes, kappa, inflate = flat_info(lib, objs)
    assert ok == 1 and nc <= 32 and ncube + ns <= 56
    sph = np.flatnonzero(objs["type"] == 1)
    assert ncube == int((objs["type"] == 2).sum())
    assert len(cull) == 9 * nc + ns                           # 8 slots + 1 bank-conflict pad per cluster, then the singles
    real = slots[slots != 255]
    assert sorted(real.tolist()) == list(range(len(sph)))     # every sphere exactly once
    assert kappa == np.float32(1.0 - 2.0 ** -18) and inflate > 0
    # every record: same centre as its sphere, inflated r^2 slightly above the exact one; every cluster box contains its spheres
    for k in range(nc):
        for j in range(8):
            s = slots[8 * k + j]
            rec = cull[9 * k + j]
            if s == 255:
                assert rec[3] < -1e29
                continue
            o = objs[sph[s]]
            r2 = np.float32(o["radius"]) * np.float32(o["radius"])
            assert np.array_equal(rec[:3], o["pos"]) and r2 < rec[3] <= r2 * 1.05 + 1e-6   # a few per cent: 4 D with D from the 3000-unit origin bound of the ground sphere
            assert np.all(boxes[k][:3] <= o["pos"] - abs(o["radius"])) and np.all(boxes[k][4:7] >= o["pos"] + abs(o["radius"]))
        assert cull[9 * k + 8][3] < -1e29                     # the pad record can never be a candidate
    for j in range(ns):
        o = objs[sph[slots[8 * nc + j]]]
        assert np.array_equal(cull[9 * nc + j][:3], o["pos"])


def test_flat_builder_limits(lib):
    rng = np.random.default_rng(3)

    def spheres(n, r=0.3):
        o = np.zeros(n, rtb200.OBJECT_DTYPE); o["type"] = 1
        o["pos"] = rng.uniform(-5, 5, (n, 3)).astype(np.float32); o["radius"] = r
        return o
    assert flat_info(lib, spheres(255))[0] == 1
    assert flat_info(lib, spheres(256))[0] == 0               # candidate codes are bytes
    assert flat_info(lib, spheres(0))[0] == 0
    o = spheres(10); o["pos"][3, 0] = np.inf
    assert flat_info(lib, o)[0] == 0                          # finite-arithmetic precondition
    o = spheres(10); o["pos"][3, 0] = 1e16
    assert flat_info(lib, o)[0] == 0
    big = spheres(200); big["radius"] = np.where(np.arange(200) % 3 == 0, 3.0, 0.05)
    assert flat_info(lib, big)[0] == 0                        # 67 level-1 singles > 56
    # margins grow with the origin extent and offset they have to cover
    a = flat_info(lib, spheres(20), extent=0.0)[4][:, 3]
    rng = np.random.default_rng(3)
    b = flat_info(lib, spheres(20), extent=1e4)[4][:, 3]
    rng = np.random.default_rng(3)
    c = flat_info(lib, spheres(20), extent=0.0, offset=50.0)[4][:, 3]
    live = a > 0
    assert np.all(b[live] > a[live]) and np.all(c[live] > a[live])


def test_obj_subset_and_triangle_records(lib, tmp_path):
    src = tmp_path / "m.obj"
    src.write_text("# comment\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0.5\nvn 0 0 1\nf 1 2 3 4\nf -4/1/1 -3/1/1 -1/1/1\n")
    verts = np.zeros((16, 3), np.float32); tris = np.zeros((16, 3), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    out = tmp_path / "m2.obj"
    r = lib.emu_obj_round_trip(str(src).encode(), str(out).encode(), p(verts), 16, p(tris), 16)
    nt, nv = r & 0xffff, r >> 16
    assert (nv, nt) == (4, 3)                                 # the quad is fanned into two triangles; negative indices are relative
    assert tris[:3].tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 3]]
    v2 = np.zeros((16, 3), np.float32); t2 = np.zeros((16, 3), np.int32)
    assert lib.emu_obj_round_trip(str(out).encode(), None, p(v2), 16, p(t2), 16) == r
    assert np.array_equal(v2, verts) and np.array_equal(t2, tris)
    assert lib.emu_obj_round_trip(str(tmp_path / "missing.obj").encode(), None, p(v2), 16, p(t2), 16) == -1
    (tmp_path / "bad.obj").write_text("v 0 0 0\nf 1 2 3\n")
    assert lib.emu_obj_round_trip(str(tmp_path / "bad.obj").encode(), None, p(v2), 16, p(t2), 16) == -1   # index out of range
    # triangle records against an independent double-precision evaluation
    rng = np.random.default_rng(5)
    v = rng.uniform(-3, 3, (30, 3)).astype(np.float32); t = rng.integers(0, 30, (40, 3)).astype(np.int32)
    t[7] = [2, 2, 5]                                          # degenerate
    pos = np.array([0.5, -1.0, 2.0], np.float32)
    rec = np.zeros((40, 12), np.float32); bounds = np.zeros((40, 6), np.float32)
    assert lib.emu_tri_records(p(pos), p(v), 30, p(t), 40, p(rec), p(bounds)) == 40
    w = (v + pos).astype(np.float32).astype(np.float64)
    for i in range(40):
        a, b, c = w[t[i]]
        n = np.cross(b - a, c - a)
        if not n.any():
            assert not rec[i].any()
            continue
        nn = n / np.linalg.norm(n)
        assert np.allclose(rec[i][:3], nn, rtol=0, atol=1e-7) and abs(rec[i][3] - nn @ a) < 1e-5
        for (px, u_want, v_want) in ((a, 0, 0), (b, 1, 0), (c, 0, 1), ((a + b + c) / 3, 1 / 3, 1 / 3)):
            u = rec[i][4:7].astype(np.float64) @ px + rec[i][7]; vv = rec[i][8:11].astype(np.float64) @ px + rec[i][11]
            assert abs(u - u_want) < 1e-4 * max(1, np.abs(rec[i][4:7]).max() * 3) and abs(vv - v_want) < 1e-4 * max(1, np.abs(rec[i][8:11]).max() * 3)
        assert np.allclose(bounds[i][:3], w[t[i]].min(0)) and np.allclose(bounds[i][3:], w[t[i]].max(0))


# ---- 8-wide quantised BVH (csrc/bvh_wide.h): every primitive inside the decoded box of the leaf child that holds it ----
def wide_info(lib, objs, extent=0.0):
    objs = np.ascontiguousarray(objs, rtb200.OBJECT_DTYPE)
    counts = (C.c_int * 5)()
    bad = lib.emu_wide_info(objs.ctypes.data_as(C.c_void_p), len(objs), C.c_float(extent), counts)
    return bad, list(counts)


@pytest.mark.parametrize("scene", SCENES)
def test_wide_bvh_builder_on_bundled_scenes(lib, scenes, scene):
    bad, (usable, n_nodes, n_refs, depth, n_b2) = wide_info(lib, scenes[scene], extent=10.0)
    prims = int(np.isin(scenes[scene]["type"], (1, 2)).sum())
    assert usable == 1 and bad == 0 and n_refs == prims and 1 <= depth <= 8 and n_nodes <= n_b2


def test_wide_bvh_builder_limits(lib):
    rng = np.random.default_rng(11)

    def spheres(n, spread=30.0):
        o = np.zeros(n, rtb200.OBJECT_DTYPE); o["type"] = 1
        o["pos"] = rng.uniform(-spread, spread, (n, 3)).astype(np.float32); o["radius"] = rng.uniform(0.05, 1.0, n).astype(np.float32)
        return o
    for n in (1, 2, 9, 64, 65, 4000):
        o = spheres(n)
        o["type"][::5] = 2; o["half"][::5] = rng.uniform(0.1, 2.0, (len(o["half"][::5]), 3)).astype(np.float32)
        bad, (usable, n_nodes, n_refs, depth, n_b2) = wide_info(lib, o, extent=100.0)
        assert usable == 1 and bad == 0 and n_refs == n, n
        assert n_nodes * 2 <= max(n_b2, 2) or n < 9, (n, n_nodes, n_b2)       # the collapse really merges levels
    # duplicates and flat (zero-extent) axes quantise without losing anything
    o = spheres(200); o["pos"][:, 1] = 2.0; o["pos"][50:100] = o["pos"][0]
    bad, (usable, *_rest) = wide_info(lib, o)
    assert usable == 1 and bad == 0
    # coordinates beyond 1e12 (or non-finite geometry, which the BVH2 keeps as a 1e30 box): not usable, callers keep the binary tree
    o = spheres(50); o["pos"][7, 0] = 3e12
    assert wide_info(lib, o)[1][0] == 0
    o = spheres(50); o["pos"][7, 2] = np.inf
    assert wide_info(lib, o)[1][0] == 0
    assert wide_info(lib, spheres(0) if False else np.zeros(0, rtb200.OBJECT_DTYPE))[0] == 0     # empty scene


# ---- BVH2 builder: concurrent subtree builds give the same bytes as the sequential build -------------------------------------
def bvh_digest(lib, objs, threads, mesh=None):
    objs = np.ascontiguousarray(objs, rtb200.OBJECT_DTYPE)
    out = (C.c_ulonglong * 4)()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    if mesh is None:
        lib.emu_bvh_digest(p(objs), len(objs), None, 0, None, 0, -1, threads, out)
    else:
        v = np.ascontiguousarray(mesh[0], np.float32).reshape(-1, 3); t = np.ascontiguousarray(mesh[1], np.int32).reshape(-1, 3)
        lib.emu_bvh_digest(p(objs), len(objs), p(v), len(v), p(t), len(t), 0, threads, out)
    return tuple(out)


def test_bvh_parallel_build_is_byte_identical(lib):
    from rtb200.scenes import heightfield_mesh, mesh_scene, synthetic_spheres
    objs = synthetic_spheres(60000, cubes_every=7)                    # well above the 16 384-primitive task threshold
    want = bvh_digest(lib, objs, 1)
    assert want[2] > 20000 and 10 < want[3] < 40
    for threads in (2, 3, 8, 64, 0):
        assert bvh_digest(lib, objs, threads) == want, threads
    v, tr = heightfield_mesh(320, 200, seed=4)                         # 128 000 triangles
    want = bvh_digest(lib, mesh_scene(), 1, (v, tr))
    assert want[2] > 40000
    for threads in (4, 16, 0):
        assert bvh_digest(lib, mesh_scene(), threads, (v, tr)) == want, threads
    small = synthetic_spheres(300)                                     # below the threshold: the sequential path whatever the count
    assert bvh_digest(lib, small, 8) == bvh_digest(lib, small, 1)


def test_quantised_nodes_enclose_the_float_boxes(lib):
    """bvh_build.h HostQNodes: every decoded 16-bit plane lies at least one grid unit outside the float plane it replaces (the
    decode's own rounding is below 0.6 units), the links are copied verbatim, and the boxes grow little on the benchmark
    scenes; a scene whose extent dwarfs its primitives is refused (the float nodes are traversed then)."""
    from rtb200.scenes import synthetic_spheres, heightfield_mesh, mesh_scene
    lib.emu_qnodes_check.restype = None
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None

    def check(objs, mesh=None, extent=30.0):
        objs = np.ascontiguousarray(objs, rtb200.OBJECT_DTYPE)
        out = (C.c_double * 4)()
        mv = np.ascontiguousarray(mesh[0], np.float32) if mesh else None
        mt = np.ascontiguousarray(mesh[1], np.int32) if mesh else None
        lib.emu_qnodes_check(p(objs), len(objs), p(mv), len(mv) if mesh else 0, p(mt), len(mt) if mesh else 0, 0 if mesh else -1, C.c_float(extent), out)
        return list(out)
    usable, slack, growth, n = check(synthetic_spheres(10000))            # BASELINE config 3 (ground sphere r = 1000: a 2000-unit grid)
    assert usable == 1 and 1.0 <= slack < 2.01 and 1.0 < growth < 1.35 and n > 2000, (usable, slack, growth, n)
    v, tr = heightfield_mesh(256, 128)
    usable, slack, growth, n = check(mesh_scene(), (v, tr))
    assert usable == 1 and 1.0 <= slack < 2.01 and 1.0 < growth < 1.25, (usable, slack, growth)
    tiny = synthetic_spheres(3000)
    tiny["radius"][:3000] *= 0.002                                         # 3000 specks in a 2000-unit scene: the grid is far too coarse
    assert check(tiny)[0] == 0
