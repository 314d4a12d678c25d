import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "software-raytracer_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
SCENES = ["Scene1", "Scene1_reflection", "Scene2", "Scene3", "Scene3_indirect", "Scene_indirect"]
SEED = (0x1234ABCD, 0x0BADC0DE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="session")
def meta():
    with open(os.path.join(GOLD, "meta.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLD, name + ".npz"))
        return cache[name]
    return load


@pytest.fixture(scope="session")
def scenes(golden):
    z = golden("bundled_scenes")
    return {s: z[s] for s in SCENES}


@pytest.fixture(scope="session")
def oracle():
    from oracle_py import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own compiled code; only where oracle/_ref was built (not on the GPU box
    unless the .so travelled). Tests using it skip otherwise."""
    from oracle_py import Reference, REFERENCE_DIR
    if not Reference.available() or not os.path.isdir(REFERENCE_DIR):
        pytest.skip("oracle/_ref or /root/reference not available here")
    return Reference()


def make_camera(cls, meta=None, rotated=False):
    """default camera (Raytracer.cpp:295-297) or the golden 'rotated' pose, as a ctypes struct of class cls."""
    c = cls()
    if rotated:
        r = meta["rotated_camera"]
        for k in ("pos", "right", "up", "forward"):
            for i in range(3):
                getattr(c, k)[i] = r[k][i]
        c.fov_deg = r["fov_deg"]
    else:
        c.right[0] = 1; c.up[1] = 1; c.forward[2] = 1; c.fov_deg = 55
    return c


@pytest.fixture(scope="session")
def tracer():
    """One PathTracer on cuda:0 for the gpu-marked tests. Fails loudly (no fallback)."""
    import rtb200
    t = rtb200.PathTracer(0)
    yield t
    t.close()
