// oracle/_ref harness: drives the REFERENCE'S OWN hot-path code, compiled from the sources
// where they lie under $REF_DIR (never copied into this repo), behind a small extern "C"
// surface that tests/ and oracle/make_goldens.py bind with ctypes.
//
// TEST INFRASTRUCTURE ONLY. Nothing in the product library links, loads or calls this.
//
// What is the reference's and what is the harness's:
//   * Common.hpp, Object.hpp, Scene.hpp                      - included as they are.
//   * Raytracer.cpp:30-59, 61-213, 223-257                   - extracted by line range at build
//     time into oracle/_ref/*.inc (git-ignored) and included textually below. Line 60
//     (`Color colorBuffer[H*W]`, a compile-time sized array) is the one declaration the
//     harness supplies itself so resolution can be chosen at run time.
//   * main() (Raytracer.cpp:259-615) is SDL/ImGui/Win32 and is not compiled; the few lines of
//     it that matter to the hot path (SunDirection normalisation :264, the strip split
//     :330-341, the frame gate :374-384 and the accumulation state machine :572-595) are
//     re-driven by ref_render_frames() below.
#include <atomic>
#include <mutex>

#include "Common.hpp"
#include "Object.hpp"
#include "Scene.hpp"

int g_ref_w = 640, g_ref_h = 480;

#include "rt_030_059.inc"
Color* colorBuffer = nullptr;  // Raytracer.cpp:60, run-time sized
#include "rt_061_213.inc"
bool suspendAllThreads = false;               // Raytracer.cpp:218
bool* threadGroupStatus = new bool[THREADS];  // Raytracer.cpp:220
#include "rt_223_257.inc"

#undef rand

// ---------------------------------------------------------------------------------------
// rand() provider (see shim.h). Lives in rng.cpp (compiled without the shim).
extern "C" {
void ref_rng_mode(int mode);                   // 0 = MSVC LCG (thread-local), 1 = Philox, 2 = glibc rand()
void ref_rng_seed_lcg(uint32_t s);
void ref_rng_key(uint32_t k0, uint32_t k1);
void ref_rng_begin_path(uint32_t pixel, uint32_t sample);
}

namespace {
Scene* g_scene = nullptr;
Transform g_camera;
std::vector<uint32_t> g_surface_pixels;
SDL_Surface g_surface;
SDL_PixelFormat g_format;
bool g_sun_normalized = false;
std::atomic<long long> g_segment_calls{0};
thread_local long long t_segment_calls = 0;    // per-thread, folded into the global when a worker exits

// a never-hit Object appended last; calls / 1 = closest-hit queries.
class CountingObject : public Object {
public:
    Rayhit Raytrace(const float3&, const float3&) const override {
        ++t_segment_calls;                     // no shared atomic on the hot path: do not slow the baseline
        return Rayhit();
    }
};
CountingObject* g_counter = nullptr;

void ensure_init() {
    if (!g_sun_normalized) {
        SunDirection = SunDirection.Normalized();  // Raytracer.cpp:264
        g_sun_normalized = true;
        progressiveResolutionScaler = 1;           // Raytracer.cpp:271
    }
}
int index_of(const Object* o) {
    for (size_t i = 0; i < ObjectsToRender.size(); ++i)
        if (ObjectsToRender[i] == o) return (int)i;
    return -1;
}
}  // namespace

extern "C" {

int ref_load_scene(const char* path) {
    ensure_init();
    if (g_scene) { g_scene->Unload(); delete g_scene; g_scene = nullptr; }
    g_counter = nullptr;
    selectedObject = NULL;
    g_scene = new Scene(std::string(path));
    g_scene->Load();                               // Raytracer.cpp:291-292
    ObjectsToRender = g_scene->GetObjects();       // Raytracer.cpp:293
    return (int)ObjectsToRender.size();
}

// Material/geometry as the reference loaded them; 16 floats per object:
// type(0 none,1 sphere,2 cube), pos3, radius|0, half3, base3, emissive3, spec3  -> then smooth, specAmt
int ref_get_objects(float* out19, int max_objects) {
    int n = 0;
    for (Object* o : ObjectsToRender) {
        if (o == g_counter) continue;
        if (n >= max_objects) break;
        float* r = out19 + 19 * n;
        Sphere* s = dynamic_cast<Sphere*>(o);
        Box* b = dynamic_cast<Box*>(o);
        r[0] = s ? 1.f : (b ? 2.f : 0.f);
        r[1] = o->transform.position.x; r[2] = o->transform.position.y; r[3] = o->transform.position.z;
        r[4] = s ? s->GetRadius() : 0.f;
        r[5] = b ? b->size.x : 0.f; r[6] = b ? b->size.y : 0.f; r[7] = b ? b->size.z : 0.f;
        r[8] = o->material.BaseColor.r; r[9] = o->material.BaseColor.g; r[10] = o->material.BaseColor.b;
        r[11] = o->material.EmissiveColor.r; r[12] = o->material.EmissiveColor.g; r[13] = o->material.EmissiveColor.b;
        r[14] = o->material.SpecularColor.r; r[15] = o->material.SpecularColor.g; r[16] = o->material.SpecularColor.b;
        r[17] = o->material.Smoothness; r[18] = o->material.SpecularAmount;
        ++n;
    }
    return n;
}

int ref_save_scene(const char* path) {
    if (!g_scene) return -1;
    g_scene->SaveAs(std::string(path));            // Scene.hpp:88-104
    return 0;
}

void ref_set_resolution(int w, int h) {
    ensure_init();
    g_ref_w = w; g_ref_h = h;
    delete[] colorBuffer;
    colorBuffer = new Color[(size_t)w * h];
    g_surface_pixels.assign((size_t)w * h, 0u);
    g_format.BytesPerPixel = 4;
    g_surface.pixels = g_surface_pixels.data();
    g_surface.pitch = w * 4;
    g_surface.format = &g_format;
    renderSurface = &g_surface;
}

void ref_set_params(int fov, int max_bounces, int simple_draw, float screen_scale) {
    ensure_init();
    FOV = fov; MAXBOUNCES = max_bounces; SIMPLEDRAW = simple_draw != 0; SCREEN_SCALE = screen_scale;
}

void ref_set_camera(const float* pos, const float* right, const float* up, const float* fwd) {
    g_camera.position = float3(pos[0], pos[1], pos[2]);
    g_camera.right = float3(right[0], right[1], right[2]);
    g_camera.up = float3(up[0], up[1], up[2]);
    g_camera.forward = float3(fwd[0], fwd[1], fwd[2]);
}

// Transform::RotateAboutAxis as the viewer applies it to the camera (Raytracer.cpp:394-395).
void ref_rotate_camera(float angle, const float* axis) {
    g_camera.RotateAboutAxis(angle, float3(axis[0], axis[1], axis[2]));
}
void ref_get_camera(float* out12) {
    const float3* v[4] = {&g_camera.position, &g_camera.right, &g_camera.up, &g_camera.forward};
    for (int i = 0; i < 4; ++i) { out12[3*i] = v[i]->x; out12[3*i+1] = v[i]->y; out12[3*i+2] = v[i]->z; }
}

void ref_select_object(int index) {
    selectedObject = (index >= 0 && index < (int)ObjectsToRender.size()) ? ObjectsToRender[index] : NULL;
}

void ref_get_env_constants(float* out16) {
    out16[0] = SunDirection.x; out16[1] = SunDirection.y; out16[2] = SunDirection.z; out16[3] = 0;
    const Color* c[3] = {&SkyColor, &HorizonColor, &GroundColor};
    for (int i = 0; i < 3; ++i) { out16[4+3*i] = c[i]->r; out16[5+3*i] = c[i]->g; out16[6+3*i] = c[i]->b; }
    out16[13] = SunColor.r; out16[14] = SunColor.g; out16[15] = SunColor.b;
}

// GetRayDirection (Raytracer.cpp:106-122) for every pixel, y-up row-major.
void ref_ray_dirs(float* out_xyz) {
    for (int y = 0; y < g_ref_h; ++y)
        for (int x = 0; x < g_ref_w; ++x) {
            float3 d = GetRayDirection(g_camera, x, y);
            float* o = out_xyz + 3 * ((size_t)x + (size_t)y * g_ref_w);
            o[0] = d.x; o[1] = d.y; o[2] = d.z;
        }
}

// Primary visibility AOVs through GetRayDirection + GetClosestObject (Raytracer.cpp:123-140).
void ref_primary_aov(int32_t* id, float* t, float* normal, float* point) {
    for (int y = 0; y < g_ref_h; ++y)
        for (int x = 0; x < g_ref_w; ++x) {
            size_t p = (size_t)x + (size_t)y * g_ref_w;
            float3 d = GetRayDirection(g_camera, x, y);
            RayHitObject h = GetClosestObject(g_camera.position, d);
            if (h.rayHit.valid) {
                id[p] = index_of(h.objectReference);
                t[p] = h.rayHit.distance;
                normal[3*p] = h.rayHit.normal.x; normal[3*p+1] = h.rayHit.normal.y; normal[3*p+2] = h.rayHit.normal.z;
                if (point) { point[3*p] = h.rayHit.point.x; point[3*p+1] = h.rayHit.point.y; point[3*p+2] = h.rayHit.point.z; }
            } else {
                id[p] = -1; t[p] = 0.f;
                normal[3*p] = normal[3*p+1] = normal[3*p+2] = 0.f;
                if (point) point[3*p] = point[3*p+1] = point[3*p+2] = 0.f;
            }
        }
}

// Arbitrary rays through GetClosestObject (secondary-ray / quirk parity).
void ref_trace_rays(const float* origin, const float* dir, int n, int32_t* id, float* t, float* normal, float* point) {
    for (int i = 0; i < n; ++i) {
        RayHitObject h = GetClosestObject(float3(origin[3*i], origin[3*i+1], origin[3*i+2]),
                                          float3(dir[3*i], dir[3*i+1], dir[3*i+2]));
        if (h.rayHit.valid) {
            id[i] = index_of(h.objectReference);
            t[i] = h.rayHit.distance;
            normal[3*i] = h.rayHit.normal.x; normal[3*i+1] = h.rayHit.normal.y; normal[3*i+2] = h.rayHit.normal.z;
            point[3*i] = h.rayHit.point.x; point[3*i+1] = h.rayHit.point.y; point[3*i+2] = h.rayHit.point.z;
        } else {
            id[i] = -1; t[i] = 0.f;
            for (int k = 0; k < 3; ++k) { normal[3*i+k] = 0.f; point[3*i+k] = 0.f; }
        }
    }
}

void ref_env_color(const float* dir, int n, float* out_rgb) {
    for (int i = 0; i < n; ++i) {
        Color c = GetEnvironmentColor(float3(dir[3*i], dir[3*i+1], dir[3*i+2]));
        out_rgb[3*i] = c.r; out_rgb[3*i+1] = c.g; out_rgb[3*i+2] = c.b;
    }
}

// RaytraceScene (Raytracer.cpp:141-213) per pixel and sample with rand() fed from the Philox
// stream keyed (pixel, sample): out_sum[p] = sum over s in [s0, s0+n) of the returned Color,
// added in sample order as plain floats. Optional per-sample dump (n*W*H*3 floats).
void ref_render_philox(uint32_t seed_lo, uint32_t seed_hi, int s0, int n, float* out_sum_rgb, float* out_per_sample_rgb) {
    ref_rng_mode(1);
    ref_rng_key(seed_lo, seed_hi);
    for (int y = 0; y < g_ref_h; ++y)
        for (int x = 0; x < g_ref_w; ++x) {
            size_t p = (size_t)x + (size_t)y * g_ref_w;
            float3 d = GetRayDirection(g_camera, x, y);
            float sr = 0.f, sg = 0.f, sb = 0.f;
            for (int s = 0; s < n; ++s) {
                ref_rng_begin_path((uint32_t)p, (uint32_t)(s0 + s));
                Color c = RaytraceScene(g_camera.position, d);
                sr += c.r; sg += c.g; sb += c.b;
                if (out_per_sample_rgb) {
                    float* o = out_per_sample_rgb + 3 * ((size_t)s * g_ref_w * g_ref_h + p);
                    o[0] = c.r; o[1] = c.g; o[2] = c.b;
                }
            }
            out_sum_rgb[3*p] = sr; out_sum_rgb[3*p+1] = sg; out_sum_rgb[3*p+2] = sb;
        }
}

// SetScreenPixel (Raytracer.cpp:63-76): feed explicit colours through accumulate + Reinhard +
// ARGB8 pack. Returns the surface (y-down rows) and the accumulation buffer (y-up).
void ref_set_pixels(const float* rgba, int set_frame, int accumulation_frames) {
    setFrame = set_frame != 0;
    ACCUMULATIONFRAMES = accumulation_frames;
    for (int y = 0; y < g_ref_h; ++y)
        for (int x = 0; x < g_ref_w; ++x) {
            const float* c = rgba + 4 * ((size_t)x + (size_t)y * g_ref_w);
            Color col; col.r = c[0]; col.g = c[1]; col.b = c[2]; col.a = c[3];  // bypass the clamping ctor: raw lanes
            SetScreenPixel(x, y, col);
        }
}
void ref_get_surface(uint32_t* out) { memcpy(out, g_surface_pixels.data(), g_surface_pixels.size() * 4); }
void ref_get_color_buffer(float* out_rgba) {
    for (size_t i = 0; i < (size_t)g_ref_w * g_ref_h; ++i) {
        out_rgba[4*i] = colorBuffer[i].r; out_rgba[4*i+1] = colorBuffer[i].g;
        out_rgba[4*i+2] = colorBuffer[i].b; out_rgba[4*i+3] = colorBuffer[i].a;
    }
}

void ref_count_segments(int enable) {
    if (enable && !g_counter) {
        g_counter = new CountingObject();
        ObjectsToRender.push_back(g_counter);
    } else if (!enable && g_counter) {
        ObjectsToRender.pop_back();
        delete g_counter; g_counter = nullptr;
    }
    g_segment_calls = 0; t_segment_calls = 0;
}
long long ref_segments() { long long v = g_segment_calls.load() + t_segment_calls; return v; }

// The reference's own frame loop: THREADS persistent workers over column strips
// (Raytracer.cpp:330-342), released and gated per frame the way main() does (:374-384,
// :572-595). rng_mode 0 = per-thread MSVC LCG (what the shipped Windows binary does),
// 2 = glibc's global locked rand(). first frame overwrites (setFrame), the rest accumulate.
// Returns wall seconds over the frame loop only.
static float g_prog_scaler = 1.0f;     // progressiveResolutionScaler for the next ref_render_frames (Raytracer.cpp:233,580)
void ref_set_progressive_scaler(float s) { g_prog_scaler = s; }
static uint32_t g_thread_seed = 1u;   // MSVC: every new thread's rand() state starts at 1
void ref_set_thread_seed(uint32_t s) { g_thread_seed = s; }
double ref_render_frames(int frames, int rng_mode, int start_frame) {
    ref_rng_mode(rng_mode);
    suspendAllThreads = false;
    progressiveResolutionScaler = g_prog_scaler;
    std::vector<std::thread*> workers(THREADS);
    volatile bool* status = threadGroupStatus;
    for (int i = 0; i < THREADS; ++i) status[i] = true;   // parked until the first release
    int div = (int)ceil(SCREEN_WIDTH / (THREADS)) + 1;     // Raytracer.cpp:330 (integer division first)
    for (int i = 0; i < THREADS; ++i) {
        int initialX = div * i;
        int nextX = min(initialX + div, SCREEN_WIDTH);
        if (initialX > SCREEN_WIDTH) initialX = SCREEN_WIDTH;
        workers[i] = new std::thread([=]() {
            ref_rng_mode(rng_mode);
            ref_rng_seed_lcg(g_thread_seed == 1u ? 1u : g_thread_seed + 7919u * (uint32_t)i);
            t_segment_calls = 0;
            renderArea(i, initialX, nextX, 0, SCREEN_HEIGHT, &g_camera);
            g_segment_calls.fetch_add(t_segment_calls, std::memory_order_relaxed);
        });
    }
    auto t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; ++f) {
        int frame_no = start_frame + f;                    // 1-based ACCUMULATIONFRAMES
        setFrame = (frame_no == 1);
        ACCUMULATIONFRAMES = frame_no;
        std::atomic_thread_fence(std::memory_order_seq_cst);
        for (int i = 0; i < THREADS; ++i) status[i] = false;   // Raytracer.cpp:592-595
        for (;;) {                                              // Raytracer.cpp:374-384
            bool busy = false;
            for (int i = 0; i < THREADS; ++i) if (!status[i]) { busy = true; break; }
            if (!busy) break;
            std::this_thread::yield();
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    suspendAllThreads = true;                              // Raytracer.cpp:598-602
    for (int i = 0; i < THREADS; ++i) { workers[i]->join(); delete workers[i]; }
    return std::chrono::duration<double>(t1 - t0).count();
}

// Mouse picking (Raytracer.cpp:530-541): x, y in window space (y-down).
int ref_pick(int x, int y_window) {
    int y = SCREEN_HEIGHT - y_window;
    RayHitObject h = GetClosestObject(g_camera.position, GetRayDirection(g_camera, x, y));
    return h.rayHit.valid ? index_of(h.objectReference) : -1;
}

}  // extern "C"
