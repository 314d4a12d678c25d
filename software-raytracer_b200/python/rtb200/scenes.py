"""Synthetic workloads of BASELINE.json's configs."""
import numpy as np

from . import OBJECT_DTYPE


def synthetic_spheres(n=10000, seed=12345, cubes_every=0):
    """Config 3: n random spheres (centres uniform in [-50,50]x[0.2,20]x[5,105], radii U[0.2,1]) over a ground
    sphere r=1000 with 8 emissive spheres r=3 E=30 at y=30; materials by thirds: diffuse, metal
    (SpecAmt 1, Smooth U[0.6,1], SpecColor=Base), 'dielectric-like' (SpecAmt 0.1, Smooth 1, SpecColor 1).
    numpy default_rng(seed) (PCG64)."""
    rng = np.random.default_rng(seed)
    o = np.zeros(n + 9, OBJECT_DTYPE)
    o["type"] = 1
    o["pos"][:n] = rng.uniform([-50, 0.2, 5], [50, 20, 105], (n, 3)).astype(np.float32)
    o["radius"][:n] = rng.uniform(0.2, 1.0, n).astype(np.float32)
    o["base"][:n] = rng.uniform(0.1, 0.95, (n, 3)).astype(np.float32)
    o["spec_color"] = 1
    third = n // 3
    o["spec_amount"][third:2 * third] = 1
    o["smoothness"][third:2 * third] = rng.uniform(0.6, 1, third).astype(np.float32)
    o["spec_color"][third:2 * third] = o["base"][third:2 * third]
    o["spec_amount"][2 * third:n] = 0.1
    o["smoothness"][2 * third:n] = 1
    o["pos"][n] = [0, -1000, 50]; o["radius"][n] = 1000; o["base"][n] = 0.8
    for k in range(8):
        o["pos"][n + 1 + k] = [-45 + 12.5 * k, 30, 20 + 10 * k]; o["radius"][n + 1 + k] = 3
        o["emissive"][n + 1 + k] = 30; o["base"][n + 1 + k] = 1
    if cubes_every:
        idx = np.arange(0, n, cubes_every)
        o["type"][idx] = 2
        o["half"][idx] = rng.uniform(0.2, 0.8, (len(idx), 3)).astype(np.float32)
    return o


def config3_camera(default_camera):
    cam = default_camera(60)
    cam.pos[0], cam.pos[1], cam.pos[2] = 0.0, 8.0, -20.0
    return cam
