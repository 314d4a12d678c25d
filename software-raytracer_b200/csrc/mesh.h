// mesh.h - triangle meshes: an EXTENSION of the reference's object model (BASELINE.json config 4).
//
// The reference has two primitives (Sphere, Box: Object.hpp:85-233) and no triangle type, so nothing
// here is pinned by reference behaviour ("parity unpinned"). What is kept from the
// reference is the Object::Raytrace CONTRACT (Object.hpp:21-23, Common.hpp:320-325): a primitive reports
// Rayhit{valid, normal, point, distance}, the closest hit wins with a strict '<' in list order
// (Raytracer.cpp:127-137), a mesh is ONE scene object (one id, one Material, Transform.position as a
// translation) and its triangles are tested in index order.
//
// The intersector is plane-first (not Moeller-Trumbore) so that it composes with a conservative BVH:
//   t = (dn - n.o) / (n.d)            reject |n.d| < 1e-9, t outside [1e-4, 10000]
//   P = o + d*t                        the reported point
//   u = m1.P + k1,  v = m2.P + k2      reject unless u >= 0, v >= 0, u + v <= 1
//   normal = n facing the ray (n.d < 0)
// Every accepted P lies on the ray and inside the triangle up to float rounding of a few coordinate
// ulps, i.e. inside the triangle's inflated box whatever the grazing angle. The 12 floats per triangle
// (n, dn, m1, k1, m2, k2) are computed here in double precision from the float vertices and rounded
// once; tests/ recompute them independently in the oracle.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"

namespace rtb {

struct HostMesh {
    std::vector<float> vertices;       // xyz, object space
    std::vector<int32_t> indices;      // 3 per triangle
    std::string file;                  // OBJ path it came from ("" = inline / API)
    bool empty() const { return indices.empty(); }
};

struct TriRecords {
    std::vector<float> rec;            // 12 floats per triangle: (n, dn) (m1, k1) (m2, k2)
    std::vector<float> bounds;         // 6 floats per triangle: lo.xyz hi.xyz of the world-space vertices
    std::vector<int32_t> obj;          // object id per triangle
    int count() const { return (int)obj.size(); }
};

// World-space vertex = object-space vertex + Transform.position, one float add per component.
// Triangles with an out-of-range index are dropped.
void build_tri_records(const std::vector<rt_object>& objects, const std::vector<HostMesh>& meshes, TriRecords& out);

// Wavefront OBJ subset: "v x y z" and "f a b c [d ...]" (fans; a/b/c index forms; negative indices).
// Returns false with a message on I/O or syntax errors.
bool load_obj(const std::string& path, HostMesh& out, std::string& err);
bool save_obj(const std::string& path, const HostMesh& mesh, std::string& err);

}  // namespace rtb
