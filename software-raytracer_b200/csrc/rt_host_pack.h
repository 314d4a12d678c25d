// rt_host_pack.h - host-side packing shared by the C-ABI (rt_capi.cu) and the host emulation used
// by the CPU tests: the pixel-independent half of GetRayDirection and the AoS -> SoA scene layout.
#pragma once
#include <cmath>
#include <vector>

#include "../../include/rt_b200.h"
#include "rt_device.cuh"

namespace rtb {

// Host half of GetRayDirection (Raytracer.cpp:107-117) and the environment constants
// (Raytracer.cpp:55-59,79,82): everything that does not depend on the pixel, in the reference's
// expression order, so device code never calls tanf.
inline void fill_frame_view(const rt_camera& cam, const rt_params& p, FrameView& f) {
    auto v3 = [](const float* v) { return make_float3(v[0], v[1], v[2]); };
    const float clip = .01f;
    float aspect = (float)p.width / (float)p.height;
    float hFov = cam.fov_deg * M_PI / 180.0f;                 // int * double / float -> double -> float (:112)
    float rd = (clip * tanf(hFov / 2.0f)) * aspect;           // :114
    float ld = (clip * tanf(hFov / 2.0f));                    // :115
    f.cam_pos = v3(cam.pos);
    f.u_axis = make_float3(cam.right[0] * rd, cam.right[1] * rd, cam.right[2] * rd);
    f.v_axis = make_float3(cam.up[0] * ld, cam.up[1] * ld, cam.up[2] * ld);
    f.fwd = make_float3(cam.forward[0] * clip, cam.forward[1] * clip, cam.forward[2] * clip);
    f.sun_neg = make_float3(p.sun_dir[0] * -1, p.sun_dir[1] * -1, p.sun_dir[2] * -1);
    float thr = (float)0.99;                                  // (double)dot > 0.99  <=>  dot >= thr
    if (!((double)thr > 0.99)) thr = nextafterf(thr, INFINITY);
    f.sun_thr = thr;
    auto clamp0 = [](float v) { return v < 0 ? 0.f : v; };
    f.sky = v3(p.sky); f.horizon = v3(p.horizon); f.ground = v3(p.ground); f.sun = v3(p.sun);
    f.sky10 = make_float3(clamp0(p.sky[0] * 0.1f), clamp0(p.sky[1] * 0.1f), clamp0(p.sky[2] * 0.1f));
    f.dissipation = p.dissipation; f.eps = p.eps;
    f.width = p.width; f.height = p.height; f.max_bounces = p.max_bounces; f.mode = p.mode; f.selected_id = p.selected_id;
    f.seed_lo = p.seed_lo; f.seed_hi = p.seed_hi;
    for (uint32_t r = 0; r < 10; ++r) { f.key_sched[2 * r] = p.seed_lo + r * 0x9E3779B9u; f.key_sched[2 * r + 1] = p.seed_hi + r * 0xBB67AE85u; }
}

// Scene object list -> dense per-type geometry lists + per-object material records.
inline void pack_scene(const std::vector<rt_object>& objs, std::vector<float4>& sph, std::vector<int>& sph_id,
                       std::vector<float4>& box, std::vector<int>& box_id, std::vector<float4>& mat) {
    sph.clear(); sph_id.clear(); box.clear(); box_id.clear(); mat.clear();
    mat.reserve(objs.size() * 3);
    for (size_t i = 0; i < objs.size(); ++i) {
        const rt_object& o = objs[i];
        if (o.type == RT_OBJ_SPHERE) {
            sph.push_back(make_float4(o.pos[0], o.pos[1], o.pos[2], o.radius * o.radius));   // squaredRadius Object.hpp:122
            sph_id.push_back((int)i);
        } else if (o.type == RT_OBJ_CUBE) {
            box.push_back(make_float4(o.pos[0], o.pos[1], o.pos[2], 0.f));
            box.push_back(make_float4(o.half[0], o.half[1], o.half[2], 0.f));
            box_id.push_back((int)i);
        }
        mat.push_back(make_float4(o.base[0], o.base[1], o.base[2], o.smoothness));
        mat.push_back(make_float4(o.emissive[0], o.emissive[1], o.emissive[2], o.spec_amount));
        mat.push_back(make_float4(o.spec_color[0], o.spec_color[1], o.spec_color[2], 0.f));
    }
}

}  // namespace rtb
