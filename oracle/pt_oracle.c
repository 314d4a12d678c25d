/* pt_oracle.c - plain-C, scalar, CPU restatement of the reference's path-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may load this. The product library (librt_b200.so) never links, loads or calls it and
 * has no CPU fallback.
 *
 * Parity status: PINNED. Every function below is checked bit-for-bit against the reference's
 * own code compiled headless (oracle/_ref, built from /root/reference by oracle/Makefile) on
 * all six bundled scenes: ray directions, closest-hit id/t/normal/point for primary and
 * random secondary rays, environment colours, per-sample path radiance with rand() supplied
 * from the same Philox stream, preview shading, running-mean accumulation and the ARGB8
 * resolve (tests/test_oracle_vs_ref.py; committed fixtures in tests/golden/ cover the GPU
 * box where /root/reference does not exist).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no FMA contraction, no fast-math): every
 * expression keeps the reference's operation order and rounding.  Citations are
 * file:line under /root/reference/Raytracer/.
 */
#define _GNU_SOURCE 1
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "philox.h"

/* ---- POD mirrors of the C-ABI structs (include/rt_b200.h); layout must match ---------- */
typedef struct {
    int32_t type;          /* 0 never-hit Object, 1 Sphere, 2 Cube (Scene.hpp:43-55) */
    float pos[3];          /* Transform.position */
    float radius;          /* Sphere (Object.hpp:87) */
    float half[3];         /* Box::size = half extents (Object.hpp:203) */
    float base[3];         /* Material.BaseColor (Common.hpp:297) */
    float emissive[3];     /* Material.EmissiveColor */
    float spec_color[3];   /* Material.SpecularColor */
    float smoothness;      /* Material.Smoothness */
    float spec_amount;     /* Material.SpecularAmount */
} orc_object;

typedef struct {
    float pos[3], right[3], up[3], forward[3];   /* Transform (Common.hpp:281-286) */
    int32_t fov_deg;                             /* FOV (Raytracer.cpp:31) */
} orc_camera;

typedef struct {
    int32_t width, height;                       /* SCREEN_WIDTH/HEIGHT (Raytracer.cpp:26-27) */
    int32_t max_bounces;                         /* MAXBOUNCES (Raytracer.cpp:32) */
    int32_t mode;                                /* 0 path, 1 preview/SIMPLEDRAW (Raytracer.cpp:35) */
    int32_t selected_id;                         /* selectedObject index or -1 (Raytracer.cpp:53) */
    float sun_dir[3];                            /* normalised SunDirection (Raytracer.cpp:55,264) */
    float sky[3], horizon[3], ground[3], sun[3]; /* Raytracer.cpp:56-59 */
    float dissipation;                           /* 0.8 (Raytracer.cpp:166) */
    float eps;                                   /* 1e-5 (Raytracer.cpp:177) */
    uint32_t seed_lo, seed_hi;
} orc_params;

/* ---- float3 / Color value semantics (Common.hpp:22-179, 180-280) ---------------------- */
typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }        /* :118 */
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }        /* :112 */
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }        /* :121 */
static inline v3 vdiv(v3 a, v3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }        /* :124 */
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }         /* float -> float3(s) broadcast :107 then :121 */
static inline v3 vneg(v3 a) { return V(a.x * -1, a.y * -1, a.z * -1); }                 /* :115 */
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      /* :83-85 */
static inline float vmag(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }     /* :77 */
static inline v3 vnormalized(v3 a) { float l = vmag(a); return V(a.x / l, a.y / l, a.z / l); }   /* :159-162: three divisions */
static inline float flerp(float a, float b, float t) { return a * (1 - t) + b * t; }    /* :19-21 */
static inline v3 vlerp(v3 a, v3 b, float t) { return V(flerp(a.x, b.x, t), flerp(a.y, b.y, t), flerp(a.z, b.z, t)); } /* :97 */
static inline v3 vreflect(v3 d, v3 n) { return vsub(d, vscale(n, 2 * vdot(d, n))); }    /* :163-165 */
static inline v3 vabs(v3 a) { return V(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }           /* :334 */
static inline v3 vsign(v3 t) {                                                          /* :328-333: sign(0) = 0 */
    return V(t.x != 0 ? t.x / fabsf(t.x) : 0, t.y != 0 ? t.y / fabsf(t.y) : 0, t.z != 0 ? t.z / fabsf(t.z) : 0);
}
static inline v3 vmax(v3 a, v3 o) { return V(a.x > o.x ? a.x : o.x, a.y > o.y ? a.y : o.y, a.z > o.z ? a.z : o.z); } /* :167 */
static inline v3 vstep(v3 e, v3 t) { return V(e.x <= t.x ? 1 : 0, e.y <= t.y ? 1 : 0, e.z <= t.z ? 1 : 0); }        /* :337 */
static inline float fmax2(float a, float b) { return a > b ? a : b; }                   /* :344-347 */
static inline float fmin2(float a, float b) { return a < b ? a : b; }                   /* :348-351 */

/* Color: every constructor call clamps negatives to 0 (Common.hpp:253-262). Only r,g,b are
 * carried: the alpha lane is 0 for every colour this path creates (ctor default a = 0). */
typedef struct { float r, g, b; } col;
static inline float c0(float v) { if (v < 0) v = 0; return v; }
static inline col C(float r, float g, float b) { col c = {c0(r), c0(g), c0(b)}; return c; }
static inline col cadd(col a, col b) { return C(a.r + b.r, a.g + b.g, a.b + b.b); }     /* :218 */
static inline col cmul(col a, col b) { return C(a.r * b.r, a.g * b.g, a.b * b.b); }     /* :212 */
static inline col cscale(col a, float s) { return C(a.r * s, a.g * s, a.b * s); }       /* :209 */
static inline col clerp(col a, col b, float t) {                                        /* :275-279, t NOT clamped */
    return C(a.r * (1 - t) + b.r * t, a.g * (1 - t) + b.g * t, a.b * (1 - t) + b.b * t);
}
static inline col cfrom(const float* p) { col c = {p[0], p[1], p[2]}; return c; }       /* stored colours: already clamped at load */

/* ---- raygen: GetRayDirection (Raytracer.cpp:106-122) ---------------------------------- */
typedef struct { v3 pos, u_axis, v_axis, fwd; int w, h; } raygen;

/* Per-frame invariants, computed on the host exactly as the reference writes them
 * (tanf from the double-computed half angle). The product computes the same thing in
 * rt_set_camera so device code never calls tanf. */
static raygen make_raygen(const orc_camera* cam, int w, int h) {
    raygen g;
    const float clip = .01f;
    float aspect = (float)w / (float)h;
    float hFov = cam->fov_deg * M_PI / 180.0f;                       /* :112 int*double/float -> double -> float */
    float rd = (clip * tanf(hFov / 2.0f)) * aspect;                  /* :114 */
    float ld = (clip * tanf(hFov / 2.0f));                           /* :115 */
    g.pos = V(cam->pos[0], cam->pos[1], cam->pos[2]);
    g.u_axis = vscale(V(cam->right[0], cam->right[1], cam->right[2]), rd);       /* right * rd   :116 */
    g.v_axis = vscale(V(cam->up[0], cam->up[1], cam->up[2]), ld);                /* up * ld      :117 */
    g.fwd = vscale(V(cam->forward[0], cam->forward[1], cam->forward[2]), clip);  /* forward*clip :113 */
    g.w = w; g.h = h;
    return g;
}
static v3 ray_dir(const raygen* g, int px, int py) {
    float nX = (px / (float)g->w) * 2 - 1;                           /* :109 pixel corner, no jitter */
    float nY = (py / (float)g->h) * 2 - 1;                           /* :110 */
    v3 u = vscale(g->u_axis, nX);
    v3 v = vscale(g->v_axis, nY);
    return vnormalized(vadd(vadd(u, v), g->fwd));                    /* :119 */
}

/* ---- intersectors --------------------------------------------------------------------- */
typedef struct { int valid; v3 normal, point; float distance; } hit_t;   /* Rayhit Common.hpp:320-325 */

/* Sphere::line_sphere_intersection (Object.hpp:104-141) */
static hit_t hit_sphere(const orc_object* s, v3 o, v3 d) {
    hit_t h; memset(&h, 0, sizeof h);
    v3 c = V(s->pos[0], s->pos[1], s->pos[2]);
    v3 L = vsub(c, o);                                   /* :115 */
    float tc = fabsf(vdot(L, d));                        /* :118-119 abs() quirk */
    v3 P = vadd(vscale(d, tc), o);                       /* :121 */
    float r2 = s->radius * s->radius;                    /* :122 */
    v3 Q = vsub(P, c);                                   /* :124 */
    float d2 = Q.x * Q.x + Q.y * Q.y + Q.z * Q.z;        /* :125 */
    if (d2 > r2) return h;                               /* :127 */
    float t1 = tc - sqrtf(r2 - d2);                      /* :131-133: no t>0 test, negative t is valid */
    h.distance = t1;
    h.point = vadd(o, vscale(d, t1));                    /* :136 */
    h.normal = vnormalized(vsub(h.point, c));            /* :137 */
    h.valid = 1;
    return h;
}

/* Box::Raytrace + iBox (Object.hpp:224-233, 173-200) */
static hit_t hit_box(const orc_object* b, v3 o, v3 rd) {
    hit_t h; memset(&h, 0, sizeof h);
    v3 ro = vsub(o, V(b->pos[0], b->pos[1], b->pos[2]));             /* :226 no rotation */
    v3 size = V(b->half[0], b->half[1], b->half[2]);
    const float lo = (float)0.01, hi = 10000;                        /* distBound :226 */
    v3 e8 = V((float)1e-8, (float)1e-8, (float)1e-8);
    v3 m = vdiv(vsign(rd), vmax(vabs(rd), e8));                      /* :175 */
    v3 n = vmul(m, ro);                                              /* :176 */
    v3 k = vmul(vabs(m), size);                                      /* :177 */
    v3 t1 = vsub(vneg(n), k);                                        /* :179 */
    v3 t2 = vadd(vneg(n), k);                                        /* :180 */
    float tN = fmax2(fmax2(t1.x, t1.y), t1.z);                       /* :181 */
    float tF = fmin2(fmin2(t2.x, t2.y), t2.z);                       /* :182 */
    float dist = FLT_MAX;
    v3 normal = V(0, 0, 0);
    if (tN > tF || tF <= 0.) {                                       /* :184 */
        dist = FLT_MAX;
    } else {
        /* :189/:193 normal always from t1, also when tF is returned */
        v3 nn = vmul(vmul(vneg(vsign(rd)), vstep(V(t1.y, t1.z, t1.x), t1)), vstep(V(t1.z, t1.x, t1.y), t1));
        if (tN >= lo && tN <= hi) { normal = nn; dist = tN; }
        else if (tF >= lo && tF <= hi) { normal = nn; dist = tF; }
        else dist = FLT_MAX;
    }
    h.normal = normal;
    h.point = vadd(o, vscale(rd, dist));                             /* :229 */
    h.distance = dist;
    h.valid = dist == FLT_MAX ? 0 : 1;                               /* :231 */
    return h;
}

/* ---- triangle meshes: EXTENSION, not in the reference ("parity unpinned") -----------
 * A mesh is one Object (type 3) whose Raytrace() reports the closest of its triangles, tested in index order
 * with a strict '<' - the Object contract of Object.hpp:21-23 / Raytracer.cpp:127-137. The plane-first
 * intersector and its 12 precomputed floats per triangle are an independent restatement of the product's
 * definition (software-raytracer_b200/csrc/mesh.h); nothing here is derived from reference behaviour. */
typedef struct { float n[3], dn, m1[3], k1, m2[3], k2; } tri_rec;
static tri_rec* g_tri = NULL;
static int32_t* g_tri_obj = NULL;
static int g_ntri = 0;

/* verts9: 3 world-space float vertices per triangle; tri_obj: owning object index, ascending */
void orc_set_triangles(const float* verts9, const int32_t* tri_obj, int n) {
    free(g_tri); free(g_tri_obj); g_tri = NULL; g_tri_obj = NULL; g_ntri = 0;
    if (n <= 0) return;
    g_tri = (tri_rec*)calloc((size_t)n, sizeof(tri_rec));
    g_tri_obj = (int32_t*)malloc((size_t)n * sizeof(int32_t));
    g_ntri = n;
    for (int t = 0; t < n; ++t) {
        const float* v = verts9 + (size_t)9 * t;
        g_tri_obj[t] = tri_obj[t];
        double e1[3], e2[3], N[3];
        for (int k = 0; k < 3; ++k) { e1[k] = (double)v[3 + k] - (double)v[k]; e2[k] = (double)v[6 + k] - (double)v[k]; }
        N[0] = e1[1] * e2[2] - e1[2] * e2[1]; N[1] = e1[2] * e2[0] - e1[0] * e2[2]; N[2] = e1[0] * e2[1] - e1[1] * e2[0];
        double nn = N[0] * N[0] + N[1] * N[1] + N[2] * N[2];
        if (!(nn > 0.0) || !isfinite(nn)) continue;                  /* degenerate: all zeros, never hit */
        double len = sqrt(nn);
        tri_rec* r = &g_tri[t];
        double nx = N[0] / len, ny = N[1] / len, nz = N[2] / len;
        r->n[0] = (float)nx; r->n[1] = (float)ny; r->n[2] = (float)nz;
        r->dn = (float)(nx * v[0] + ny * v[1] + nz * v[2]);
        double m1[3] = {(e2[1] * N[2] - e2[2] * N[1]) / nn, (e2[2] * N[0] - e2[0] * N[2]) / nn, (e2[0] * N[1] - e2[1] * N[0]) / nn};
        double m2[3] = {(N[1] * e1[2] - N[2] * e1[1]) / nn, (N[2] * e1[0] - N[0] * e1[2]) / nn, (N[0] * e1[1] - N[1] * e1[0]) / nn};
        for (int k = 0; k < 3; ++k) { r->m1[k] = (float)m1[k]; r->m2[k] = (float)m2[k]; }
        r->k1 = (float)-(m1[0] * v[0] + m1[1] * v[1] + m1[2] * v[2]);
        r->k2 = (float)-(m2[0] * v[0] + m2[1] * v[1] + m2[2] * v[2]);
    }
}

static hit_t hit_tri(const tri_rec* r, v3 o, v3 d) {
    hit_t h; memset(&h, 0, sizeof h);
    float denom = r->n[0] * d.x + r->n[1] * d.y + r->n[2] * d.z;
    if (fabsf(denom) < 1e-9f) return h;
    float t = (r->dn - (r->n[0] * o.x + r->n[1] * o.y + r->n[2] * o.z)) / denom;
    if (!(t >= 1e-4f && t <= 10000.f)) return h;
    float Px = o.x + d.x * t, Py = o.y + d.y * t, Pz = o.z + d.z * t;
    float u = (r->m1[0] * Px + r->m1[1] * Py + r->m1[2] * Pz) + r->k1;
    float v = (r->m2[0] * Px + r->m2[1] * Py + r->m2[2] * Pz) + r->k2;
    if (!(u >= 0.f && v >= 0.f && u + v <= 1.f)) return h;
    h.valid = 1; h.distance = t; h.point = V(Px, Py, Pz);
    h.normal = denom < 0.f ? V(r->n[0], r->n[1], r->n[2]) : V(r->n[0] * -1.f, r->n[1] * -1.f, r->n[2] * -1.f);
    return h;
}

/* Mesh::Raytrace: closest triangle of object `obj`, index order, strict < */
static hit_t hit_mesh(int obj, v3 o, v3 d) {
    hit_t best; memset(&best, 0, sizeof best);
    float shortest = INFINITY;
    for (int t = 0; t < g_ntri; ++t) {
        if (g_tri_obj[t] != obj) continue;
        hit_t h = hit_tri(&g_tri[t], o, d);
        if (h.valid && h.distance < shortest) { shortest = h.distance; best = h; }
    }
    return best;
}

/* GetClosestObject (Raytracer.cpp:123-140): strict <, list order, +inf start */
static int closest(const orc_object* objs, int n, v3 o, v3 d, hit_t* out, long long* segs) {
    int best = -1;
    float shortest = INFINITY;
    hit_t res; memset(&res, 0, sizeof res);
    if (segs) ++*segs;
    for (int i = 0; i < n; ++i) {
        hit_t h;
        if (objs[i].type == 1) h = hit_sphere(&objs[i], o, d);
        else if (objs[i].type == 2) h = hit_box(&objs[i], o, d);
        else if (objs[i].type == 3) h = hit_mesh(i, o, d);           /* extension */
        else continue;                                               /* Object::Raytrace: never valid (Object.hpp:21-23) */
        if (h.valid && h.distance < shortest) { best = i; shortest = h.distance; res = h; }
    }
    *out = res;
    return best;
}

/* ---- environment: GetEnvironmentColor (Raytracer.cpp:77-89) --------------------------- */
static col env_color(const orc_params* p, v3 d) {
    float upd = vdot(d, V(0, 1, 0));                                 /* :78 WORLDUP */
    v3 sd = V(p->sun_dir[0], p->sun_dir[1], p->sun_dir[2]);
    col sun = (vdot(d, vscale(sd, -1)) > 0.99) ? cfrom(p->sun) : C(0, 0, 0);   /* :79 float dot vs double 0.99 */
    col sky = cfrom(p->sky), hor = cfrom(p->horizon), gnd = cfrom(p->ground);
    if (upd > 0) {
        col t = clerp(hor, sky, powf(upd, 0.1f));                    /* :81 */
        t = clerp(t, cscale(sky, 0.1f), upd);                        /* :82 */
        return cadd(t, sun);
    }
    upd = fabsf(upd);
    return cadd(clerp(hor, gnd, powf(upd, .05f)), sun);              /* :87 */
}

/* ---- RNG plumbing --------------------------------------------------------------------- */
typedef struct {
    int mode;                 /* 0 MSVC LCG, 1 Philox(pixel,sample,block) */
    uint32_t lcg;
    uint32_t key[2], pixel, sample, block, widx, buf[4];
    int have;
} rng_t;

static int rng_next(rng_t* r) {                                      /* the value rand() returns */
    if (r->mode == 0) { r->lcg = r->lcg * 214013u + 2531011u; return (int)((r->lcg >> 16) & 0x7fffu); }
    if (!r->have) {
        uint32_t ctr[4] = {r->pixel, r->sample, r->block, 0u};
        philox4x32_10(ctr, r->key, r->buf);
        r->have = 1;
    }
    uint32_t w = r->buf[r->widx];
    if (r->widx == 3) { r->block++; r->widx = 0; r->have = 0; }      /* consecutive words, 4 per block */
    else r->widx++;
    return (int)(w >> 17);
}
static void rng_begin(rng_t* r, uint32_t pixel, uint32_t sample) {
    r->pixel = pixel; r->sample = sample; r->block = 0; r->widx = 0; r->have = 0;
}
#define ORC_RAND_MAX 32767
static inline float rng_unit(rng_t* r) { return (float)rng_next(r) / ORC_RAND_MAX; }   /* (float)rand() / RAND_MAX */

/* GetRandomDirection + GetRandomNormalOrientedHemisphere (Raytracer.cpp:90-105):
 * normalize(uniform cube); the x*y*z > 1 rejection can never fire. */
static v3 hemisphere_dir(rng_t* r, v3 n) {
    v3 sr;
    do {
        sr.x = (rng_unit(r) - 0.5f) * 2;
        sr.y = (rng_unit(r) - 0.5f) * 2;
        sr.z = (rng_unit(r) - 0.5f) * 2;
    } while (sr.x * sr.y * sr.z > 1);
    sr = vnormalized(sr);
    if (vdot(sr, n) < 0) sr = vscale(sr, -1);
    return sr;
}

static inline float smoothstep_f(float e0, float e1, float x) {      /* Common.hpp:352-365 */
    if (x < e0) return 0;
    if (x >= e1) return 1;
    x = (x - e0) / (e1 - e0);
    return x * x * (3 - 2 * x);
}

/* RaytraceScene (Raytracer.cpp:141-213) */
static col radiance(const orc_object* objs, int n, const orc_params* p, v3 o, v3 d, rng_t* rng, long long* segs) {
    hit_t hit;
    int id = closest(objs, n, o, d, &hit, segs);
    if (id < 0) return env_color(p, d);                              /* :143-145 */

    if (p->mode == 1) {                                              /* SIMPLEDRAW :147-160 */
        col refl = env_color(p, vreflect(d, hit.normal));
        float k = objs[id].spec_amount, s = objs[id].smoothness;
        float fresnal = 0;
        if (id == p->selected_id) {
            fresnal = 1 - vdot(vneg(hit.normal), d);
            fresnal = fmax2(fresnal, 0.0f);
            fresnal = smoothstep_f(0.0f, 0.5f, fresnal);
        }
        col a = cadd(cadd(cscale(cfrom(objs[id].base), 1 - k), cscale(cscale(refl, k), s)), cfrom(objs[id].emissive));
        return clerp(a, C(3, 3, 0), fresnal);
    }

    col incoming = cfrom(objs[id].emissive);                         /* :162 */
    col hitColor = cfrom(objs[id].base);                             /* :163 */
    v3 sray = d;
    int coin = objs[id].spec_amount >= rng_unit(rng);                /* :165 */
    for (int i = 0; i < p->max_bounces; ++i) {
        if (i != 0) hitColor = cscale(hitColor, p->dissipation);     /* :169-171 */
        v3 refl = vreflect(sray, hit.normal);                        /* :172 */
        sray = hemisphere_dir(rng, hit.normal);                      /* :174 */
        sray = vlerp(sray, refl, objs[id].smoothness * coin);        /* :175 */
        sray = vnormalized(sray);                                    /* :176 */
        v3 org = vadd(hit.point, vscale(hit.normal, p->eps));        /* :177 */
        id = closest(objs, n, org, sray, &hit, segs);
        if (id < 0) {
            incoming = cadd(incoming, cmul(env_color(p, sray), hitColor));   /* :179 */
            break;
        }
        coin = objs[id].spec_amount >= rng_unit(rng);                /* :182 */
        incoming = cadd(incoming, cmul(cfrom(objs[id].emissive), hitColor));             /* :183 */
        hitColor = cmul(hitColor, clerp(cfrom(objs[id].base), cfrom(objs[id].spec_color), (float)coin));  /* :184 */
    }
    return incoming;
}

/* ======================================================================================= */
/* exported surface (ctypes)                                                               */
/* ======================================================================================= */

void orc_default_params(orc_params* p) {
    memset(p, 0, sizeof *p);
    p->width = 1280; p->height = 720;                                /* Raytracer.cpp:26-27 */
    p->max_bounces = 2; p->mode = 1; p->selected_id = -1;            /* :32, :35, :53 */
    v3 sd = vnormalized(V(1, -1, -1));                               /* :55, :264 */
    p->sun_dir[0] = sd.x; p->sun_dir[1] = sd.y; p->sun_dir[2] = sd.z;
    col sky = cscale(C(.2, .35, 1.0f), 10.0f);                       /* :56 */
    col hor = cscale(C(1.0, 0.9f, 0.5f), 5.0f);                      /* :57 */
    col gnd = C(.08f, .06f, .03f);                                   /* :58 */
    p->sky[0] = sky.r; p->sky[1] = sky.g; p->sky[2] = sky.b;
    p->horizon[0] = hor.r; p->horizon[1] = hor.g; p->horizon[2] = hor.b;
    p->ground[0] = gnd.r; p->ground[1] = gnd.g; p->ground[2] = gnd.b;
    p->sun[0] = p->sun[1] = p->sun[2] = 500;                         /* :59 */
    p->dissipation = 0.8f; p->eps = .00001f;                         /* :166, :177 */
}

void orc_default_camera(orc_camera* c) {                             /* Raytracer.cpp:295-297, :31 */
    memset(c, 0, sizeof *c);
    c->right[0] = 1; c->up[1] = 1; c->forward[2] = 1; c->fov_deg = 55;
}

/* Transform::RotateAboutAxis (Common.hpp:287-291): Rodrigues on each basis vector. */
static v3 vcross(v3 l, v3 r) { return V(l.y * r.z - r.y * l.z, r.x * l.z - l.x * r.z, l.x * r.y - r.x * l.y); }  /* :94-96 */
static v3 rotate_axis(v3 b, float angle, v3 axis) {
    return vadd(vadd(vscale(b, cosf(angle)), vscale(vcross(axis, b), sinf(angle))),
                vscale(vscale(axis, vdot(axis, b)), 1 - cosf(angle)));
}
void orc_rotate_camera(orc_camera* c, float angle, const float* axis3) {
    v3 ax = V(axis3[0], axis3[1], axis3[2]);
    v3 f = rotate_axis(V(c->forward[0], c->forward[1], c->forward[2]), angle, ax);
    v3 u = rotate_axis(V(c->up[0], c->up[1], c->up[2]), angle, ax);
    v3 r = rotate_axis(V(c->right[0], c->right[1], c->right[2]), angle, ax);
    c->forward[0] = f.x; c->forward[1] = f.y; c->forward[2] = f.z;
    c->up[0] = u.x; c->up[1] = u.y; c->up[2] = u.z;
    c->right[0] = r.x; c->right[1] = r.y; c->right[2] = r.z;
}

/* The per-frame raygen invariants the host hands the device: u_axis, v_axis, fwd (9 floats). */
void orc_raygen_basis(const orc_camera* cam, int w, int h, float* out9) {
    raygen g = make_raygen(cam, w, h);
    out9[0] = g.u_axis.x; out9[1] = g.u_axis.y; out9[2] = g.u_axis.z;
    out9[3] = g.v_axis.x; out9[4] = g.v_axis.y; out9[5] = g.v_axis.z;
    out9[6] = g.fwd.x; out9[7] = g.fwd.y; out9[8] = g.fwd.z;
}

void orc_ray_dirs(const orc_camera* cam, int w, int h, float* out_xyz) {
    raygen g = make_raygen(cam, w, h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            v3 d = ray_dir(&g, x, y);
            float* o = out_xyz + 3 * ((size_t)x + (size_t)y * w);
            o[0] = d.x; o[1] = d.y; o[2] = d.z;
        }
}

void orc_trace_rays(const orc_object* objs, int n, const float* origin, const float* dir, int nrays,
                    int32_t* id, float* t, float* normal, float* point) {
    for (int i = 0; i < nrays; ++i) {
        hit_t h;
        int k = closest(objs, n, V(origin[3*i], origin[3*i+1], origin[3*i+2]), V(dir[3*i], dir[3*i+1], dir[3*i+2]), &h, NULL);
        id[i] = k;
        if (k >= 0) {
            t[i] = h.distance;
            normal[3*i] = h.normal.x; normal[3*i+1] = h.normal.y; normal[3*i+2] = h.normal.z;
            if (point) { point[3*i] = h.point.x; point[3*i+1] = h.point.y; point[3*i+2] = h.point.z; }
        } else {
            t[i] = 0.f;
            normal[3*i] = normal[3*i+1] = normal[3*i+2] = 0.f;
            if (point) point[3*i] = point[3*i+1] = point[3*i+2] = 0.f;
        }
    }
}

void orc_primary_aov(const orc_object* objs, int n, const orc_camera* cam, int w, int h,
                     int32_t* id, float* t, float* normal, float* point) {
    raygen g = make_raygen(cam, w, h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t p = (size_t)x + (size_t)y * w;
            v3 d = ray_dir(&g, x, y);
            float o3[3] = {g.pos.x, g.pos.y, g.pos.z}, d3[3] = {d.x, d.y, d.z};
            orc_trace_rays(objs, n, o3, d3, 1, id + p, t + p, normal + 3 * p, point ? point + 3 * p : NULL);
        }
}

void orc_env_color(const orc_params* p, const float* dir, int n, float* out_rgb) {
    for (int i = 0; i < n; ++i) {
        col c = env_color(p, V(dir[3*i], dir[3*i+1], dir[3*i+2]));
        out_rgb[3*i] = c.r; out_rgb[3*i+1] = c.g; out_rgb[3*i+2] = c.b;
    }
}

/* ---- rendering: per pixel, samples [s0, s0+nspp) summed in sample order ---------------- */
typedef struct {
    const orc_object* objs; int n; const orc_camera* cam; const orc_params* p;
    int s0, nspp, y0, y1, rng_mode; uint32_t lcg_seed;
    float* out_sum; float* out_samples; long long segs;
} band_job;

static void* band_main(void* arg) {
    band_job* j = (band_job*)arg;
    const orc_params* p = j->p;
    int w = p->width, h = p->height;
    raygen g = make_raygen(j->cam, w, h);
    rng_t rng; memset(&rng, 0, sizeof rng);
    rng.mode = j->rng_mode; rng.lcg = j->lcg_seed; rng.key[0] = p->seed_lo; rng.key[1] = p->seed_hi;
    long long segs = 0;
    for (int y = j->y0; y < j->y1; ++y)
        for (int x = 0; x < w; ++x) {
            size_t px = (size_t)x + (size_t)y * w;
            v3 d = ray_dir(&g, x, y);
            float sr = 0.f, sg = 0.f, sb = 0.f;
            for (int s = 0; s < j->nspp; ++s) {
                rng_begin(&rng, (uint32_t)px, (uint32_t)(j->s0 + s));
                col c = radiance(j->objs, j->n, p, g.pos, d, &rng, &segs);
                sr += c.r; sg += c.g; sb += c.b;
                if (j->out_samples) {
                    float* o = j->out_samples + 3 * ((size_t)s * w * h + px);
                    o[0] = c.r; o[1] = c.g; o[2] = c.b;
                }
            }
            if (j->out_sum) { j->out_sum[3*px] = sr; j->out_sum[3*px+1] = sg; j->out_sum[3*px+2] = sb; }
        }
    j->segs = segs;
    return NULL;
}

/* rng_mode 1: Philox(pixel, sample) - result independent of `threads`.
 * rng_mode 0: MSVC LCG, one state per worker seeded 1 - timing only (what the reference does). */
long long orc_render(const orc_object* objs, int n, const orc_camera* cam, const orc_params* p,
                     int s0, int nspp, int rng_mode, int threads, float* out_sum_rgb, float* out_per_sample_rgb) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads > p->height) threads = p->height;
    band_job jobs[256]; pthread_t tid[256];
    /* interleave-free bands; rows are independent */
    for (int i = 0; i < threads; ++i) {
        band_job* j = &jobs[i];
        j->objs = objs; j->n = n; j->cam = cam; j->p = p; j->s0 = s0; j->nspp = nspp; j->rng_mode = rng_mode;
        j->lcg_seed = 1u; j->out_sum = out_sum_rgb; j->out_samples = out_per_sample_rgb; j->segs = 0;
        j->y0 = (int)((long long)p->height * i / threads);
        j->y1 = (int)((long long)p->height * (i + 1) / threads);
    }
    if (threads == 1) band_main(&jobs[0]);
    else {
        for (int i = 0; i < threads; ++i) pthread_create(&tid[i], NULL, band_main, &jobs[i]);
        for (int i = 0; i < threads; ++i) pthread_join(tid[i], NULL);
    }
    long long segs = 0;
    for (int i = 0; i < threads; ++i) segs += jobs[i].segs;
    return segs;                                                     /* closest-hit queries traced */
}

/* ---- SetScreenPixel (Raytracer.cpp:63-76) + Color -> Uint32 (Common.hpp:189-206) -------- */
static inline uint32_t pack_lane(float v) {
    /* (int)(v*255): out-of-range / NaN conversions are what cvttss2si gives on the reference's
     * x86 targets (INT_MIN), then the >255 clamp and the (Uint8) truncation. */
    float s = v * 255;
    int i;
    if (!(s > -2147483904.0f && s < 2147483648.0f)) i = INT32_MIN; else i = (int)s;
    if (i > 255) i = 255;
    return (uint32_t)(uint8_t)i;
}
static inline float cdiv_lane(float a, float b) { return c0(a / b); }                   /* Color / Color :215 */

/* Reinhard + pack of one accumulated colour: finalColor / (Color(1,1,1) + finalColor).
 * Alpha: 0/(0+0) = NaN -> byte 0. */
uint32_t orc_resolve_pixel(float r, float g, float b) {
    float R = cdiv_lane(r, c0(1 + r)), G = cdiv_lane(g, c0(1 + g)), B = cdiv_lane(b, c0(1 + b));
    return (0u << 24) | (pack_lane(R) << 16) | (pack_lane(G) << 8) | pack_lane(B);
}

/* accum_rgba: W*H float4 (y-up). If `count` > 0 the buffer holds SUMS of `count` samples
 * (the product's representation) and is divided first; count == 0 means it already holds
 * the mean (the reference's representation). out: rows at (H-1-y) when flip_y. */
void orc_resolve_argb8(const float* accum_rgba, int w, int h, int count, int flip_y, uint32_t* out, int pitch_bytes) {
    for (int y = 0; y < h; ++y) {
        uint32_t* row = (uint32_t*)((uint8_t*)out + (size_t)(flip_y ? h - 1 - y : y) * pitch_bytes);
        for (int x = 0; x < w; ++x) {
            const float* a = accum_rgba + 4 * ((size_t)x + (size_t)y * w);
            float r = a[0], g = a[1], b = a[2];
            if (count > 0) { float c = (float)count; r = r / c; g = g / c; b = b / c; }
            row[x] = orc_resolve_pixel(r, g, b);
        }
    }
}

/* The reference's accumulation: running mean with weight (float)(1.0/frames) (Raytracer.cpp:65-71). */
void orc_running_mean(float* buf_rgb, const float* color_rgb, size_t n_pixels, int set_frame, int frames) {
    float weight = 1.0 / frames;
    for (size_t i = 0; i < 3 * n_pixels; ++i) {
        if (set_frame) buf_rgb[i] = color_rgb[i];
        else buf_rgb[i] = c0(c0(buf_rgb[i] * (1 - weight)) + c0(color_rgb[i] * weight));
    }
}

/* Timed CPU baseline ("port"): nspp frames of the whole image on `threads` workers.
 * Returns seconds; *segments = closest-hit queries. */
double orc_time_render(const orc_object* objs, int n, const orc_camera* cam, const orc_params* p,
                       int nspp, int rng_mode, int threads, long long* segments) {
    struct timespec a, b;
    float* sum = (float*)malloc((size_t)p->width * p->height * 3 * sizeof(float));
    clock_gettime(CLOCK_MONOTONIC, &a);
    long long s = orc_render(objs, n, cam, p, 0, nspp, rng_mode, threads, sum, NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    free(sum);
    if (segments) *segments = s;
    return (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
}

/* Philox known-answer hook for tests. */
void orc_philox(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) { philox4x32_10(ctr4, key2, out4); }
int orc_sizeof_object(void) { return (int)sizeof(orc_object); }
int orc_sizeof_params(void) { return (int)sizeof(orc_params); }
int orc_sizeof_camera(void) { return (int)sizeof(orc_camera); }
