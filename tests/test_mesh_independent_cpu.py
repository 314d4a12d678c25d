"""CPU: the oracle's triangle intersector (oracle/pt_oracle.c hit_tri, which restates the product's plane-first test) against
an INDEPENDENT float64 Moeller-Trumbore brute force on sampled rays (tests/mt_reference.py). The
reference has no triangles, so this is what pins the mesh extension's geometry from outside the product's own formula;
tests/test_gpu_round2.py runs the same check on the CUDA path."""
import numpy as np

import rtb200
from rtb200.scenes import heightfield_mesh, mesh_scene
from mt_reference import moller_trumbore, check_against_mt


def sampled_rays(n, seed):
    rng = np.random.default_rng(seed)
    org = rng.uniform([-8, 0.5, 0], [8, 4, 12], (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d[:, 1] = -np.abs(d[:, 1]) * 0.7
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d = (d / np.sqrt((d ** 2).sum(1, dtype=np.float32))[:, None]).astype(np.float32)
    # a third of the rays start below the surface and look up: back faces
    org[::3, 1] = -3.0; d[::3, 1] = np.abs(d[::3, 1])
    return org, d


def test_oracle_triangles_against_float64_moller_trumbore(oracle):
    v, tr = heightfield_mesh(96, 64, seed=21)             # 12 288 triangles
    objs = mesh_scene()[:1]                               # the mesh object alone: every hit is a triangle
    org, d = sampled_rays(1500, 8)
    oracle.set_triangles(objs, {0: (v, tr)})
    try:
        ids, t, nrm, pt = oracle.trace_rays(objs, org, d)
    finally:
        oracle.set_triangles(objs, {})
    world = v.astype(np.float64) + objs["pos"][0].astype(np.float64)
    mt = moller_trumbore(world, tr, org, d)
    n_hit, n_diff = check_against_mt(ids, t, nrm, mt)
    assert n_hit > 500 and n_diff <= 3
    # hit points lie on the ray
    hit = ids == 0
    assert np.allclose(pt[hit], org[hit].astype(np.float64) + d[hit].astype(np.float64) * t[hit, None], rtol=0, atol=2e-5)
