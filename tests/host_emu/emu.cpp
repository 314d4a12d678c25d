// CPU emulation of the render kernel's per-lane loop using the PRODUCT'S OWN device functions
// (csrc/rt_device.cuh compiled for the host through cuda_shim.h, -ffp-contract=off) and host
// code (scene packing, frame constants, BVH builder). TEST INFRASTRUCTURE: lets the CPU-only
// test run check the device-side logic against the oracle without a GPU. Never shipped, never
// loaded by the product.
#include "cuda_shim.h"
#define RTB_HOST_EMULATION 1
#include "../../software-raytracer_b200/csrc/rt_device.cuh"
#include "../../software-raytracer_b200/csrc/rt_bvh_lane.cuh"
#include "../../software-raytracer_b200/csrc/rt_host_pack.h"
#include "../../software-raytracer_b200/csrc/bvh_build.h"
#include "../../software-raytracer_b200/csrc/bvh_wide.h"
#include "../../software-raytracer_b200/csrc/flat_build.h"
#include "../../software-raytracer_b200/csrc/mesh.h"

using namespace rtb;

extern "C" {
// Sum over samples [s0, s0+n) for every pixel (float3 per pixel, y-up); accel 0 = brute force,
// 1 = BVH candidates, 2 = flat two-level accelerator, 3 = 8-wide quantised BVH (bvh_wide.h). Also returns primary AOVs when the pointers are given. Returns segments traced.
// Optional mesh (extension): object `mesh_object` (type RT_OBJ_MESH) gets the given triangles.
long long emu_render_mesh(const rt_object* objects, int n_obj, const rt_camera* cam, const rt_params* par, int accel,
                          uint32_t s0, int n, float* out_rgb, int32_t* aov_id, float* aov_t, float* aov_n, float* aov_p,
                          const float* mverts, int n_mverts, const int32_t* mtris, int n_mtris, int mesh_object);
long long emu_render(const rt_object* objects, int n_obj, const rt_camera* cam, const rt_params* par, int accel,
                     uint32_t s0, int n, float* out_rgb, int32_t* aov_id, float* aov_t, float* aov_n, float* aov_p) {
    return emu_render_mesh(objects, n_obj, cam, par, accel, s0, n, out_rgb, aov_id, aov_t, aov_n, aov_p, nullptr, 0, nullptr, 0, -1);
}
long long emu_render_mesh(const rt_object* objects, int n_obj, const rt_camera* cam, const rt_params* par, int accel,
                          uint32_t s0, int n, float* out_rgb, int32_t* aov_id, float* aov_t, float* aov_n, float* aov_p,
                          const float* mverts, int n_mverts, const int32_t* mtris, int n_mtris, int mesh_object) {
    std::vector<rt_object> objs(objects, objects + n_obj);
    std::vector<HostMesh> meshes((size_t)n_obj);
    if (mesh_object >= 0 && mesh_object < n_obj) {
        meshes[(size_t)mesh_object].vertices.assign(mverts, mverts + (size_t)3 * n_mverts);
        meshes[(size_t)mesh_object].indices.assign(mtris, mtris + (size_t)3 * n_mtris);
    }
    TriRecords tris;
    build_tri_records(objs, meshes, tris);
    std::vector<float4> sph, box, mat; std::vector<int> sph_id, box_id;
    pack_scene(objs, sph, sph_id, box, box_id, mat);
    SceneView sc;
    sc.sph = sph.data(); sc.sph_id = sph_id.data(); sc.box = box.data(); sc.box_id = box_id.data(); sc.mat = mat.data();
    sc.n_sph = (int)sph_id.size(); sc.n_box = (int)box_id.size(); sc.n_obj = n_obj;
    sc.tri = reinterpret_cast<const float4*>(tris.rec.data()); sc.tri_obj = tris.obj.data(); sc.n_tri = tris.count();
    FrameView fr;
    fill_frame_view(*cam, *par, fr);
    HostBvh bvh;
    float ext = 0.f;
    for (int k = 0; k < 3; ++k) ext = fmaxf(ext, fabsf(cam->pos[k]));
    build_bvh(objs, ext, bvh, accel == 3 ? kWideMaxLeaf : 4, &tris);
    const float4* nodes = reinterpret_cast<const float4*>(bvh.nodes.data());
    HostWideBvh wide;
    if (accel == 3) { build_wide_bvh(bvh, wide); if (!wide.usable) return -2; }
    const int wentries = wide.depth + 2;
    std::vector<int> stack((size_t)bvh.max_depth + 8 + 3 * (size_t)wentries);
    unsigned char queue[64];
    HostFlat flat;
    build_flat(objs, ext, flat);
    if (accel == 2 && (!flat.usable || sc.n_tri > 0)) return -1;
    FlatView fv;
    fv.boxes = reinterpret_cast<const float4*>(flat.boxes.data()); fv.cull = reinterpret_cast<const float4*>(flat.cull.data());
    fv.cull_slot = flat.cull_slot.data(); fv.prim_id = flat.prim_id.data();
    fv.n_clusters = flat.n_clusters; fv.n_cubes = flat.n_cubes; fv.n_singles = flat.n_singles; fv.kappa = flat.kappa;
    std::vector<float4> oct((size_t)16 * flat.n_clusters + 1);                   // the per-octant cluster boxes, as setup_trace() stages them
    for (int i = 0; i < 8 * flat.n_clusters; ++i) flat_fill_oct(fv.boxes, i >> 3, i & 7, oct.data());
    fv.oct = oct.data();
    long long segs = 0;
    for (int py = 0; py < fr.height; ++py)
        for (int px = 0; px < fr.width; ++px) {
            const uint32_t pixel = (uint32_t)px + (uint32_t)py * (uint32_t)fr.width;
            const float3 d0 = ray_dir(fr, px, py);
            auto trace = [&](float3 o, float3 d) {
                if (accel == 2) return closest_hit_flat(sc, fv, sc.sph, sc.box, queue, 1, o, d);
                if (accel == 3) return closest_hit_bvh8(sc, sc.sph, sc.box, reinterpret_cast<const uint4*>(wide.nodes.data()), wide.refs.data(),
                                                        stack.data(), 1, wentries, 0x47u, o, d);
                return accel ? closest_hit_bvh(sc, sc.sph, sc.box, nodes, bvh.refs.data(), stack.data(), 1, o, d)
                             : closest_hit(sc, sc.sph, sc.box, o, d);
            };
            if (aov_id) {
                Hit h = trace(fr.cam_pos, d0);
                aov_id[pixel] = h.id; aov_t[pixel] = h.t;
                aov_n[3 * pixel] = h.n.x; aov_n[3 * pixel + 1] = h.n.y; aov_n[3 * pixel + 2] = h.n.z;
                aov_p[3 * pixel] = h.p.x; aov_p[3 * pixel + 1] = h.p.y; aov_p[3 * pixel + 2] = h.p.z;
            }
            float3 acc = f3(0.f, 0.f, 0.f);
            float3 o = fr.cam_pos, d = d0, T = f3(0, 0, 0), L = f3(0, 0, 0);
            int s = 0, depth = 0;
            while (s < n) {
                Hit h = trace(o, d);
                ++segs;
                float3 c;
                if (shade_segment(sc, fr, h, pixel, s0 + (uint32_t)s, o, d, T, L, depth, c)) {
                    acc.x += c.x; acc.y += c.y; acc.z += c.z;
                    ++s; depth = 0; o = fr.cam_pos; d = d0;
                }
            }
            if (out_rgb) { out_rgb[3 * pixel] = acc.x; out_rgb[3 * pixel + 1] = acc.y; out_rgb[3 * pixel + 2] = acc.z; }
        }
    return segs;
}
}

// traversal statistics of the emulated wide / binary BVH loops since the last call (test diagnostics)
extern "C" void emu_bvh_stats(long long* out4) {   // six values
    out4[0] = g_wide_node_visits; out4[1] = g_wide_prim_tests; out4[2] = g_bvh2_node_visits; out4[3] = g_wide_empty_visits; out4[4] = g_wide_stale_visits; out4[5] = g_bvh2_prim_tests; g_wide_stale_visits = g_bvh2_prim_tests = 0;
    g_wide_node_visits = g_wide_prim_tests = g_bvh2_node_visits = g_wide_empty_visits = 0;
}
// Wide BVH of a scene (spheres and cubes): counts[0..4] = usable, wide nodes, refs, depth, BVH2 nodes. Returns the number of
// violations of "every primitive is referenced exactly once, by a leaf child whose DECODED box contains the primitive's own box"
// (must be 0) - checked here independently of the builder's own containment assert.
extern "C" int emu_wide_info(const rt_object* objects, int n_obj, float origin_extent, int* counts) {
    std::vector<rt_object> objs(objects, objects + n_obj);
    HostBvh b2; HostWideBvh w;
    build_bvh(objs, origin_extent, b2, kWideMaxLeaf, nullptr);
    build_wide_bvh(b2, w);
    counts[0] = w.usable; counts[1] = (int)w.nodes.size(); counts[2] = (int)w.refs.size(); counts[3] = w.depth; counts[4] = (int)b2.nodes.size();
    if (!w.usable) return 0;
    // primitive boxes in build order; leaf refs address spheres by slot (>= 0) and cubes by ~slot (bvh_build.h)
    struct PrimBox { float lo[3], hi[3]; };
    std::vector<PrimBox> sph_box, cube_box;
    for (const rt_object& o : objs) {
        if (o.type != RT_OBJ_SPHERE && o.type != RT_OBJ_CUBE) continue;
        PrimBox b;
        for (int k = 0; k < 3; ++k) {
            const float h = o.type == RT_OBJ_SPHERE ? fabsf(o.radius) : fabsf(o.half[k]);
            b.lo[k] = o.pos[k] - h; b.hi[k] = o.pos[k] + h;
        }
        (o.type == RT_OBJ_SPHERE ? sph_box : cube_box).push_back(b);
    }
    std::vector<int> sph_seen(sph_box.size(), 0), cube_seen(cube_box.size(), 0);
    int bad = 0;
    for (const WideNode& nd : w.nodes) {
        const uint8_t* hdr = reinterpret_cast<const uint8_t*>(&nd.w[3]);     // scale exponents x, y, z; imask
        const uint8_t* meta = reinterpret_cast<const uint8_t*>(&nd.w[6]);
        const uint8_t* qb = reinterpret_cast<const uint8_t*>(&nd.w[8]);      // qlo.x qlo.y qlo.z qhi.x qhi.y qhi.z, 8 each
        float org[3]; memcpy(org, &nd.w[0], 12);
        for (int s = 0; s < 8; ++s) {
            if (meta[s] == 0 || ((hdr[3] >> s) & 1)) continue;               // empty slot or inner child
            const int cnt = __builtin_popcount(meta[s] >> 5), off = meta[s] & 31;
            for (int j = 0; j < cnt; ++j) {
                const int r = w.refs[(size_t)nd.w[5] + off + j];
                const PrimBox& pb = r >= 0 ? sph_box[(size_t)r] : cube_box[(size_t)(~r)];
                ++(r >= 0 ? sph_seen[(size_t)r] : cube_seen[(size_t)(~r)]);
                for (int k = 0; k < 3; ++k) {
                    const double sc = ldexp(1.0, (int)hdr[k] - 127);
                    const double lo = (double)org[k] + sc * qb[8 * k + s], hi = (double)org[k] + sc * qb[24 + 8 * k + s];
                    if (!(lo <= pb.lo[k] && hi >= pb.hi[k])) ++bad;
                }
            }
        }
    }
    for (int c : sph_seen) if (c != 1) ++bad;
    for (int c : cube_seen) if (c != 1) ++bad;
    return bad;
}

// BVH2 of a scene (optionally with one mesh object) built with `threads` threads: out[0..3] = FNV-1a 64 of the node array, of
// the refs, node count, max depth. The parallel build must give the same bytes as the sequential one.
extern "C" void emu_bvh_digest(const rt_object* objects, int n_obj, const float* mverts, int n_mverts, const int32_t* mtris, int n_mtris,
                               int mesh_object, int threads, unsigned long long* out4) {
    std::vector<rt_object> objs(objects, objects + n_obj);
    std::vector<HostMesh> meshes((size_t)n_obj);
    if (mesh_object >= 0 && mesh_object < n_obj) {
        meshes[(size_t)mesh_object].vertices.assign(mverts, mverts + (size_t)3 * n_mverts);
        meshes[(size_t)mesh_object].indices.assign(mtris, mtris + (size_t)3 * n_mtris);
    }
    TriRecords tris;
    build_tri_records(objs, meshes, tris);
    HostBvh b;
    build_bvh(objs, 25.f, b, 4, &tris, 1e-5f, threads);
    auto fnv = [](const void* p, size_t n) {
        unsigned long long h = 1469598103934665603ull;
        const unsigned char* c = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; }
        return h;
    };
    out4[0] = fnv(b.nodes.data(), b.nodes.size() * sizeof(BvhNode));
    out4[1] = fnv(b.refs.data(), b.refs.size() * sizeof(int32_t));
    out4[2] = b.nodes.size(); out4[3] = (unsigned long long)b.max_depth;
}

// ---- host-logic hooks for the CPU tests (builders only, no tracing) ----------------------------------------------
extern "C" {
// Flat accelerator of a scene: returns usable (0/1); counts[0..3] = clusters, cubes, singles, cull records;
// cull (4 floats per record, up to max_records) and cull_slot (up to max_slots) copied out; kappa and inflation returned.
int emu_flat_info(const rt_object* objects, int n_obj, float origin_extent, float origin_offset, int* counts, float* cull,
                  int max_records, unsigned char* cull_slot, int max_slots, float* boxes, int max_boxes, float* kappa_inflate) {
    std::vector<rt_object> objs(objects, objects + n_obj);
    HostFlat f;
    build_flat(objs, origin_extent, f, origin_offset);
    counts[0] = f.n_clusters; counts[1] = f.n_cubes; counts[2] = f.n_singles; counts[3] = (int)(f.cull.size() / 4);
    for (size_t i = 0; i < f.cull.size() && i < (size_t)4 * max_records; ++i) cull[i] = f.cull[i];
    for (size_t i = 0; i < f.cull_slot.size() && i < (size_t)max_slots; ++i) cull_slot[i] = f.cull_slot[i];
    for (size_t i = 0; i < f.boxes.size() && i < (size_t)8 * max_boxes; ++i) boxes[i] = f.boxes[i];
    kappa_inflate[0] = f.kappa; kappa_inflate[1] = f.inflate_abs;
    return f.usable ? 1 : 0;
}
// OBJ subset reader/writer and the triangle records (mesh.h). Returns triangles (or -1 on a load error).
int emu_obj_round_trip(const char* path_in, const char* path_out, float* verts, int max_verts, int32_t* tris, int max_tris) {
    HostMesh m; std::string err;
    if (!load_obj(path_in, m, err)) return -1;
    if (path_out && !save_obj(path_out, m, err)) return -2;
    for (size_t i = 0; i < m.vertices.size() && i < (size_t)3 * max_verts; ++i) verts[i] = m.vertices[i];
    for (size_t i = 0; i < m.indices.size() && i < (size_t)3 * max_tris; ++i) tris[i] = m.indices[i];
    return (int)(m.indices.size() / 3) | ((int)(m.vertices.size() / 3) << 16);
}
int emu_tri_records(const float* pos3, const float* verts, int n_verts, const int32_t* tris, int n_tris, float* rec12, float* bounds6) {
    std::vector<rt_object> objs(1);
    memset(&objs[0], 0, sizeof(rt_object));
    objs[0].type = RT_OBJ_MESH; objs[0].pos[0] = pos3[0]; objs[0].pos[1] = pos3[1]; objs[0].pos[2] = pos3[2];
    std::vector<HostMesh> meshes(1);
    meshes[0].vertices.assign(verts, verts + (size_t)3 * n_verts);
    meshes[0].indices.assign(tris, tris + (size_t)3 * n_tris);
    TriRecords t;
    build_tri_records(objs, meshes, t);
    for (size_t i = 0; i < t.rec.size(); ++i) rec12[i] = t.rec[i];
    for (size_t i = 0; i < t.bounds.size(); ++i) bounds6[i] = t.bounds[i];
    return t.count();
}
}


// The per-lane traversal state machine of the persistent kernels (csrc/rt_bvh_lane.cuh: BvhLane + the sentinel stack) driven on
// the CPU with a RANDOM schedule of node phases and leaf phases - the device ends a node phase by a warp vote - against the plain
// per-ray loop closest_hit_bvh() on the same rays. Returns the number of rays whose hit differs in any bit (must be 0);
// *steps receives the node and leaf steps taken.
extern "C" int emu_lane_schedules(const rt_object* objects, int n_obj, const float* mverts, int n_mverts, const int32_t* mtris, int n_mtris, int mesh_object,
                                  const float* org, const float* dir, int n_rays, unsigned seed, int leaf_bias, int use_slots, long long* steps) {
    std::vector<rt_object> objs(objects, objects + n_obj);
    std::vector<HostMesh> meshes((size_t)n_obj);
    if (mesh_object >= 0 && mesh_object < n_obj) {
        meshes[(size_t)mesh_object].vertices.assign(mverts, mverts + (size_t)3 * n_mverts);
        meshes[(size_t)mesh_object].indices.assign(mtris, mtris + (size_t)3 * n_mtris);
    }
    TriRecords tris;
    build_tri_records(objs, meshes, tris);
    std::vector<float4> sph, box, mat; std::vector<int> sph_id, box_id;
    pack_scene(objs, sph, sph_id, box, box_id, mat);
    SceneView sc;
    sc.sph = sph.data(); sc.sph_id = sph_id.data(); sc.box = box.data(); sc.box_id = box_id.data(); sc.mat = mat.data();
    sc.n_sph = (int)sph_id.size(); sc.n_box = (int)box_id.size(); sc.n_obj = n_obj;
    sc.tri = reinterpret_cast<const float4*>(tris.rec.data()); sc.tri_obj = tris.obj.data(); sc.n_tri = tris.count();
    float ext = 0.f;
    for (int i = 0; i < 3 * n_rays; ++i) ext = fmaxf(ext, fabsf(org[i]));
    HostBvh bvh;
    build_bvh(objs, ext, bvh, 4, &tris);
    if (bvh.max_depth + 2 > 62) return -1;
    const float4* nodes = reinterpret_cast<const float4*>(bvh.nodes.data());
    std::vector<int> stack((size_t)bvh.max_depth + 8);
    std::vector<float> slots;                                // leaf-ordered primitive slots (bvh_build.h), what MODE 3 kernels read
    if (use_slots) build_leaf_slots(bvh, reinterpret_cast<const float*>(sph.data()), sph_id.data(), reinterpret_cast<const float*>(box.data()), box_id.data(), &tris, slots);
    const float4* slot4 = use_slots ? reinterpret_cast<const float4*>(slots.data()) : nullptr;
    // use_slots == 2: also the 32-byte quantised nodes (bvh_build.h HostQNodes) - the conservative encoding AND the device's decode
    HostQNodes qn;
    if (use_slots == 2) { build_qnodes(bvh, qn); if (!qn.usable) return -2; }
    const uint4* qnodes = use_slots == 2 ? reinterpret_cast<const uint4*>(qn.words.data()) : nullptr;
    const float3 q_org = f3(qn.org[0], qn.org[1], qn.org[2]), q_step = f3(qn.step[0], qn.step[1], qn.step[2]);
    unsigned rng = seed * 2654435761u + 12345u;
    int bad = 0;
    long long n_node = 0, n_leaf = 0;
    BvhLane L;
    L.init(nullptr);
    for (int i = 0; i < n_rays; ++i) {
        const float3 o = f3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), d = f3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        const Hit want = closest_hit_bvh(sc, sc.sph, sc.box, nodes, bvh.refs.data(), stack.data(), 1, o, d);
        TravCount cnt = {0u, 0u, 0u, 0u};
        L.begin(o, d, qnodes != nullptr, q_org, q_step);       // the stack must be empty again after every traversal
        while (L.state == BvhLane::ACTIVE) {                 // one warp iteration: a node phase of at least one step, then the leaf phase
            for (;;) {
                if (L.in_node()) { L.node_step<true, true>(nodes, cnt, qnodes, 0x4B00u); ++n_node; }
                rng = rng * 1664525u + 1013904223u;
                // leaf_bias 0: the node phase runs until the lane has no inner node left; 8: one node step per leaf phase
                if (!L.in_node() || (int)((rng >> 16) % 8u) < leaf_bias) break;
            }
            if (L.state == BvhLane::ACTIVE) { L.leaf_step<true>(sc, sc.sph, sc.box, bvh.refs.data(), slot4, o, d, cnt); ++n_leaf; }
        }
        if (L.state != BvhLane::DONE || L.ls.sp != LaneStack::kEntry) { ++bad; L.init(nullptr); continue; }
        const Hit got = L.finish(sc.sph, o, d);
        if (got.id != want.id || memcmp(&got.t, &want.t, 4) || memcmp(&got.n, &want.n, 12) || memcmp(&got.p, &want.p, 12)) ++bad;
    }
    if (steps) { steps[0] = n_node; steps[1] = n_leaf; }
    return bad;
}


// Quantised nodes of a scene's BVH (bvh_build.h HostQNodes), checked independently of the builder's own assert: out[0] = usable,
// out[1] = smallest slack (in grid units) between a decoded plane and the float plane it must enclose (>= 1 by construction),
// out[2] = mean ratio decoded / float box surface area over all child boxes, out[3] = nodes.
extern "C" void emu_qnodes_check(const rt_object* objects, int n_obj, const float* mverts, int n_mverts, const int32_t* mtris, int n_mtris, int mesh_object,
                                 float origin_extent, double* out) {
    std::vector<rt_object> objs(objects, objects + n_obj);
    std::vector<HostMesh> meshes((size_t)n_obj);
    if (mesh_object >= 0 && mesh_object < n_obj) {
        meshes[(size_t)mesh_object].vertices.assign(mverts, mverts + (size_t)3 * n_mverts);
        meshes[(size_t)mesh_object].indices.assign(mtris, mtris + (size_t)3 * n_mtris);
    }
    TriRecords tris;
    build_tri_records(objs, meshes, tris);
    HostBvh bvh;
    build_bvh(objs, origin_extent, bvh, 4, &tris);
    HostQNodes qn;
    build_qnodes(bvh, qn);
    out[0] = qn.usable ? 1 : 0; out[1] = 1e300; out[2] = 0; out[3] = (double)bvh.nodes.size();
    if (!qn.usable) return;
    double ratio = 0; long long boxes = 0;
    for (size_t i = 0; i < bvh.nodes.size(); ++i) {
        const BvhNode& nd = bvh.nodes[i];
        for (int c = 0; c < 2; ++c) {
            double e[3], q[3]; bool empty = false;
            for (int k = 0; k < 3; ++k) {
                const double a = nd.f[6 * c + 2 * k], b = nd.f[6 * c + 2 * k + 1];
                if (!(a <= b)) { empty = true; break; }
                const uint32_t w = qn.words[8 * i + 3 * c + k];
                const double lo = (double)qn.org[k] + (double)(w & 0xffffu) * (double)qn.step[k], hi = (double)qn.org[k] + (double)(w >> 16) * (double)qn.step[k];
                out[1] = std::min(out[1], std::min((a - lo) / (double)qn.step[k], (hi - b) / (double)qn.step[k]));
                e[k] = b - a; q[k] = hi - lo;
            }
            if (empty) continue;
            const double sa = e[0] * e[1] + e[1] * e[2] + e[2] * e[0], sq = q[0] * q[1] + q[1] * q[2] + q[2] * q[0];
            if (sa > 0) { ratio += sq / sa; ++boxes; }
        }
        if ((int32_t)qn.words[8 * i + 6] != nd.c[0] || (int32_t)qn.words[8 * i + 7] != nd.c[1]) out[1] = -1;     // links must be copied verbatim
    }
    out[2] = boxes ? ratio / (double)boxes : 0;
}
