"""ctypes bindings for the CPU oracle (liboracle.so) and, when built, the reference's own
code (oracle/_ref/libref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs. The product (software-raytracer_b200/) never imports this module.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref.so")
REFERENCE_DIR = "/root/reference/Raytracer"


class OrcObject(C.Structure):
    _fields_ = [("type", C.c_int32), ("pos", C.c_float * 3), ("radius", C.c_float), ("half", C.c_float * 3),
                ("base", C.c_float * 3), ("emissive", C.c_float * 3), ("spec_color", C.c_float * 3),
                ("smoothness", C.c_float), ("spec_amount", C.c_float)]


class OrcCamera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("right", C.c_float * 3), ("up", C.c_float * 3), ("forward", C.c_float * 3),
                ("fov_deg", C.c_int32)]


class OrcParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("max_bounces", C.c_int32), ("mode", C.c_int32),
                ("selected_id", C.c_int32), ("sun_dir", C.c_float * 3), ("sky", C.c_float * 3),
                ("horizon", C.c_float * 3), ("ground", C.c_float * 3), ("sun", C.c_float * 3),
                ("dissipation", C.c_float), ("eps", C.c_float), ("seed_lo", C.c_uint32), ("seed_hi", C.c_uint32)]


OBJECT_DTYPE = np.dtype([("type", "<i4"), ("pos", "<f4", 3), ("radius", "<f4"), ("half", "<f4", 3),
                         ("base", "<f4", 3), ("emissive", "<f4", 3), ("spec_color", "<f4", 3),
                         ("smoothness", "<f4"), ("spec_amount", "<f4")])
assert OBJECT_DTYPE.itemsize == C.sizeof(OrcObject) == 76


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def build_oracle(force=False):
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "pt_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return ORACLE_SO


def build_ref():
    """Compile the reference's own sources (only possible where /root/reference exists)."""
    if os.path.exists(os.path.join(REFERENCE_DIR, "Raytracer.cpp")):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return REF_SO if os.path.exists(REF_SO) else None


def load_scene_json(path):
    """Scene JSON -> OBJECT_DTYPE array with the defaults of Scene.hpp:59-69 / Common.hpp:313-318.
    (Test-side loader for the oracle; the product has its own C++ loader.)"""
    with open(path) as f:
        data = json.load(f)
    objs = data["SceneObjects"]
    out = np.zeros(len(objs), OBJECT_DTYPE)
    for i, o in enumerate(objs):
        r = o["Renderer"]
        rec = out[i]
        rec["pos"] = np.array(o["Position"], np.float64).astype(np.float32)
        if r["Type"] == "Sphere":
            rec["type"] = 1
            rec["radius"] = np.float32(r["Radius"])
        elif r["Type"] == "Cube":
            rec["type"] = 2
            rec["half"] = np.array(r["Size"], np.float64).astype(np.float32)
        else:
            rec["type"] = 0
        smooth, amount = 0.5, 0.0
        base, spec, emis = [1, 1, 1], [1, 1, 1], [0, 0, 0]
        if "Material" in o:
            m = o["Material"]
            smooth = m.get("Smoothness", 0.5)
            amount = m.get("SpecularAmount", 0.1)
            spec = m.get("SpecularColor", [1, 1, 1])
            base = m.get("Color", [1, 1, 1])
            emis = m.get("Emissive", [0, 0, 0])
        clamp = lambda v: np.maximum(np.array(v, np.float64).astype(np.float32), np.float32(0))
        rec["smoothness"] = np.float32(smooth)
        rec["spec_amount"] = np.float32(amount)
        rec["base"], rec["spec_color"], rec["emissive"] = clamp(base), clamp(spec), clamp(emis)
    return out


class Oracle:
    """The plain-C restatement."""

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.orc_render.restype = C.c_longlong
        L.orc_time_render.restype = C.c_double
        L.orc_resolve_pixel.restype = C.c_uint32
        L.orc_resolve_pixel.argtypes = [C.c_float] * 3
        assert L.orc_sizeof_object() == C.sizeof(OrcObject)
        assert L.orc_sizeof_params() == C.sizeof(OrcParams)
        assert L.orc_sizeof_camera() == C.sizeof(OrcCamera)

    def default_params(self, **kw):
        p = OrcParams()
        self.lib.orc_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def default_camera(self, fov=55):
        c = OrcCamera()
        self.lib.orc_default_camera(C.byref(c))
        c.fov_deg = fov
        return c

    def rotate_camera(self, cam, angle, axis):
        ax = np.asarray(axis, np.float32)
        self.lib.orc_rotate_camera(C.byref(cam), C.c_float(angle), _p(ax))

    def raygen_basis(self, cam, w, h):
        out = np.zeros(9, np.float32)
        self.lib.orc_raygen_basis(C.byref(cam), w, h, _p(out))
        return out

    def ray_dirs(self, cam, w, h):
        out = np.zeros((h, w, 3), np.float32)
        self.lib.orc_ray_dirs(C.byref(cam), w, h, _p(out))
        return out

    def set_triangles(self, objs, meshes):
        """mesh extension: meshes = {object_index: (vertices float32 (n,3) object space, triangles int32 (m,3))}.
        World-space vertex = vertex + pos with one float32 add (what the product does); pass {} to clear."""
        verts, owner = [], []
        for oi in sorted(meshes):
            v, t = meshes[oi]
            v = (np.ascontiguousarray(v, np.float32).reshape(-1, 3) + np.asarray(objs["pos"][oi], np.float32)).astype(np.float32)
            t = np.ascontiguousarray(t, np.int32).reshape(-1, 3)
            verts.append(v[t].reshape(-1, 9)); owner.append(np.full(len(t), oi, np.int32))
        if verts:
            v9 = np.ascontiguousarray(np.concatenate(verts), np.float32); ow = np.ascontiguousarray(np.concatenate(owner), np.int32)
            self.lib.orc_set_triangles(_p(v9), _p(ow), len(ow))
        else:
            self.lib.orc_set_triangles(None, None, 0)

    def trace_rays(self, objs, origin, direction):
        origin = np.ascontiguousarray(origin, np.float32)
        direction = np.ascontiguousarray(direction, np.float32)
        n = origin.shape[0]
        ids = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
        nrm = np.zeros((n, 3), np.float32); pt = np.zeros((n, 3), np.float32)
        self.lib.orc_trace_rays(_p(objs), len(objs), _p(origin), _p(direction), n, _p(ids), _p(t), _p(nrm), _p(pt))
        return ids, t, nrm, pt

    def primary_aov(self, objs, cam, w, h):
        ids = np.zeros((h, w), np.int32); t = np.zeros((h, w), np.float32)
        nrm = np.zeros((h, w, 3), np.float32); pt = np.zeros((h, w, 3), np.float32)
        self.lib.orc_primary_aov(_p(objs), len(objs), C.byref(cam), w, h, _p(ids), _p(t), _p(nrm), _p(pt))
        return ids, t, nrm, pt

    def env_color(self, params, dirs):
        dirs = np.ascontiguousarray(dirs, np.float32)
        out = np.zeros_like(dirs)
        self.lib.orc_env_color(C.byref(params), _p(dirs), dirs.shape[0], _p(out))
        return out

    def render(self, objs, cam, params, s0, nspp, rng_mode=1, threads=None, per_sample=False):
        """Sum over samples [s0, s0+nspp) per pixel, (H, W, 3) float32 y-up; returns (sum, samples|None, segments)."""
        w, h = params.width, params.height
        out = np.zeros((h, w, 3), np.float32)
        samples = np.zeros((nspp, h, w, 3), np.float32) if per_sample else None
        threads = threads or (os.cpu_count() or 1)
        segs = self.lib.orc_render(_p(objs), len(objs), C.byref(cam), C.byref(params), s0, nspp, rng_mode, threads,
                                   _p(out), _p(samples))
        return out, samples, segs

    def time_render(self, objs, cam, params, nspp, rng_mode=0, threads=None):
        segs = C.c_longlong(0)
        threads = threads or (os.cpu_count() or 1)
        sec = self.lib.orc_time_render(_p(objs), len(objs), C.byref(cam), C.byref(params), nspp, rng_mode, threads,
                                       C.byref(segs))
        return sec, segs.value

    def resolve_argb8(self, accum_rgba, count, flip_y=True):
        accum_rgba = np.ascontiguousarray(accum_rgba, np.float32)
        h, w = accum_rgba.shape[:2]
        out = np.zeros((h, w), np.uint32)
        self.lib.orc_resolve_argb8(_p(accum_rgba), w, h, count, int(flip_y), _p(out), w * 4)
        return out

    def running_mean(self, buf_rgb, color_rgb, set_frame, frames):
        self.lib.orc_running_mean(_p(buf_rgb), _p(color_rgb), C.c_size_t(buf_rgb.size // 3), int(set_frame), frames)

    def philox(self, ctr, key):
        ctr = np.asarray(ctr, np.uint32); key = np.asarray(key, np.uint32); out = np.zeros(4, np.uint32)
        self.lib.orc_philox(_p(ctr), _p(key), _p(out))
        return out


class Reference:
    """The reference's own compiled code (oracle/_ref). Exists only where it was built."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.lib = C.CDLL(REF_SO)
        self.lib.ref_render_frames.restype = C.c_double
        self.lib.ref_segments.restype = C.c_longlong
        self.w = self.h = 0

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def load_scene(self, path):
        n = self.lib.ref_load_scene(str(path).encode())
        return n

    def objects(self):
        buf = np.zeros((4096, 19), np.float32)
        n = self.lib.ref_get_objects(_p(buf), 4096)
        out = np.zeros(n, OBJECT_DTYPE)
        out["type"] = buf[:n, 0].astype(np.int32)
        out["pos"] = buf[:n, 1:4]; out["radius"] = buf[:n, 4]; out["half"] = buf[:n, 5:8]
        out["base"] = buf[:n, 8:11]; out["emissive"] = buf[:n, 11:14]; out["spec_color"] = buf[:n, 14:17]
        out["smoothness"] = buf[:n, 17]; out["spec_amount"] = buf[:n, 18]
        return out

    def save_scene(self, path):
        return self.lib.ref_save_scene(str(path).encode())

    def setup(self, w, h, fov=55, max_bounces=8, simple_draw=False, cam=None, screen_scale=1.0):
        self.w, self.h = w, h
        self.lib.ref_set_resolution(w, h)
        self.lib.ref_set_params(fov, max_bounces, int(simple_draw), C.c_float(screen_scale))
        if cam is None:
            vecs = ([0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1])
        else:
            vecs = (list(cam.pos), list(cam.right), list(cam.up), list(cam.forward))
        arrs = [np.array(v, np.float32) for v in vecs]
        self.lib.ref_set_camera(*[_p(a) for a in arrs])

    def rotate_camera(self, angle, axis):
        ax = np.asarray(axis, np.float32)
        self.lib.ref_rotate_camera(C.c_float(angle), _p(ax))

    def camera(self):
        out = np.zeros(12, np.float32)
        self.lib.ref_get_camera(_p(out))
        return out.reshape(4, 3)

    def select(self, idx):
        self.lib.ref_select_object(idx)

    def env_constants(self):
        out = np.zeros(16, np.float32)
        self.lib.ref_get_env_constants(_p(out))
        return out

    def ray_dirs(self):
        out = np.zeros((self.h, self.w, 3), np.float32)
        self.lib.ref_ray_dirs(_p(out))
        return out

    def primary_aov(self):
        h, w = self.h, self.w
        ids = np.zeros((h, w), np.int32); t = np.zeros((h, w), np.float32)
        nrm = np.zeros((h, w, 3), np.float32); pt = np.zeros((h, w, 3), np.float32)
        self.lib.ref_primary_aov(_p(ids), _p(t), _p(nrm), _p(pt))
        return ids, t, nrm, pt

    def trace_rays(self, origin, direction):
        origin = np.ascontiguousarray(origin, np.float32)
        direction = np.ascontiguousarray(direction, np.float32)
        n = origin.shape[0]
        ids = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
        nrm = np.zeros((n, 3), np.float32); pt = np.zeros((n, 3), np.float32)
        self.lib.ref_trace_rays(_p(origin), _p(direction), n, _p(ids), _p(t), _p(nrm), _p(pt))
        return ids, t, nrm, pt

    def env_color(self, dirs):
        dirs = np.ascontiguousarray(dirs, np.float32)
        out = np.zeros_like(dirs)
        self.lib.ref_env_color(_p(dirs), dirs.shape[0], _p(out))
        return out

    def render_philox(self, seed_lo, seed_hi, s0, nspp, per_sample=False):
        out = np.zeros((self.h, self.w, 3), np.float32)
        samples = np.zeros((nspp, self.h, self.w, 3), np.float32) if per_sample else None
        self.lib.ref_render_philox(C.c_uint32(seed_lo), C.c_uint32(seed_hi), s0, nspp, _p(out), _p(samples))
        return out, samples

    def set_pixels(self, rgba, set_frame, frames):
        rgba = np.ascontiguousarray(rgba, np.float32)
        self.lib.ref_set_pixels(_p(rgba), int(set_frame), frames)
        surf = np.zeros((self.h, self.w), np.uint32)
        self.lib.ref_get_surface(_p(surf))
        buf = np.zeros((self.h, self.w, 4), np.float32)
        self.lib.ref_get_color_buffer(_p(buf))
        return surf, buf

    def render_frames(self, frames, rng_mode=0, start_frame=1, count_segments=False, thread_seed=1, progressive_scaler=1.0):
        """The reference's own 16-thread frame loop; returns (seconds, segments|None).
        thread_seed 1 = MSVC semantics (every worker's rand() starts at seed 1)."""
        self.lib.ref_set_thread_seed(C.c_uint32(thread_seed))
        self.lib.ref_set_progressive_scaler(C.c_float(progressive_scaler))
        self.lib.ref_count_segments(int(count_segments))
        sec = self.lib.ref_render_frames(frames, rng_mode, start_frame)
        segs = self.lib.ref_segments() if count_segments else None
        if count_segments:
            self.lib.ref_count_segments(0)
        return sec, segs

    def color_buffer(self):
        buf = np.zeros((self.h, self.w, 4), np.float32)
        self.lib.ref_get_color_buffer(_p(buf))
        return buf

    def surface(self):
        surf = np.zeros((self.h, self.w), np.uint32)
        self.lib.ref_get_surface(_p(surf))
        return surf

    def pick(self, x, y_window):
        return self.lib.ref_pick(x, y_window)
