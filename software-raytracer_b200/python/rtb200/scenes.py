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


def heightfield_mesh(nx=1024, nz=512, seed=12345, size=(20.0, 10.0), amplitude=0.6):
    """Config 4: tessellated, displaced heightfield: (nx x nz) quads -> 2*nx*nz triangles (1024 x 512 -> 1.05 M),
    object space: x in [-size0/2, size0/2], z in [0, size1], y = smooth pseudo-random displacement.
    Returns (vertices float32 (n,3), triangles int32 (m,3))."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-size[0] / 2, size[0] / 2, nx + 1, dtype=np.float64)
    zs = np.linspace(0, size[1], nz + 1, dtype=np.float64)
    X, Z = np.meshgrid(xs, zs)
    Y = np.zeros_like(X)
    for _ in range(6):                                  # a few random sine waves: smooth hills
        kx, kz = rng.uniform(0.3, 2.5, 2); ph = rng.uniform(0, 2 * np.pi, 2); a = rng.uniform(0.2, 1.0)
        Y += a * np.sin(kx * X + ph[0]) * np.cos(kz * Z + ph[1])
    Y *= amplitude / 3.0
    v = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float32)
    i = (np.arange(nz)[:, None] * (nx + 1) + np.arange(nx)[None, :]).reshape(-1)
    t = np.concatenate([np.stack([i, i + nx + 1, i + 1], -1), np.stack([i + 1, i + nx + 1, i + nx + 2], -1)]).astype(np.int32)
    return v, t


def mesh_scene():
    """Object list of config 4: the mesh object (index 0, diffuse-ish), an emissive sphere and two reflective spheres
    above it. Attach heightfield_mesh() to object 0 with PathTracer.set_mesh."""
    o = np.zeros(4, OBJECT_DTYPE)
    o["spec_color"] = 1
    o["type"][0] = 3; o["pos"][0] = [0, -1.0, 3.0]; o["base"][0] = [0.75, 0.7, 0.6]; o["smoothness"][0] = 0.3; o["spec_amount"][0] = 0.1
    o["type"][1] = 1; o["pos"][1] = [3.0, 5.0, 9.0]; o["radius"][1] = 1.5; o["emissive"][1] = 40; o["base"][1] = 1
    o["type"][2] = 1; o["pos"][2] = [-1.5, 0.4, 6.0]; o["radius"][2] = 0.9; o["base"][2] = [0.9, 0.3, 0.3]; o["smoothness"][2] = 1; o["spec_amount"][2] = 0.6
    o["type"][3] = 1; o["pos"][3] = [1.8, 0.2, 7.5]; o["radius"][3] = 0.8; o["base"][3] = [0.3, 0.5, 0.9]; o["smoothness"][3] = 0.9; o["spec_amount"][3] = 1
    return o
