// mesh.cpp - triangle records and the OBJ subset reader/writer. See mesh.h.
#include "mesh.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <system_error>
#include <fstream>
#include <sstream>

namespace rtb {

namespace {
// one triangle of mesh m (object oi) -> 12-float record, bounds, object id at output slot `slot`
void tri_record(const rt_object& o, const HostMesh& m, size_t t, size_t oi, size_t slot, TriRecords& out) {
    const int ia = m.indices[t], ib = m.indices[t + 1], ic = m.indices[t + 2];
    float v[3][3];
    const int idx[3] = {ia, ib, ic};
    for (int c = 0; c < 3; ++c)
        for (int k = 0; k < 3; ++k) v[c][k] = m.vertices[(size_t)3 * idx[c] + k] + o.pos[k];   // float add
    double e1[3], e2[3], N[3];
    for (int k = 0; k < 3; ++k) { e1[k] = (double)v[1][k] - (double)v[0][k]; e2[k] = (double)v[2][k] - (double)v[0][k]; }
    N[0] = e1[1] * e2[2] - e1[2] * e2[1]; N[1] = e1[2] * e2[0] - e1[0] * e2[2]; N[2] = e1[0] * e2[1] - e1[1] * e2[0];
    const double nn = N[0] * N[0] + N[1] * N[1] + N[2] * N[2];
    float* r = out.rec.data() + 12 * slot;
    if (!(nn > 0.0) || !std::isfinite(nn)) {
        for (int k = 0; k < 12; ++k) r[k] = 0.f;                     // degenerate: n.d == 0, never hit
    } else {
        const double len = std::sqrt(nn);
        double n[3] = {N[0] / len, N[1] / len, N[2] / len};
        // u = m1.(P - v0) with m1 = e2 x N / |N|^2 ; v = m2.(P - v0) with m2 = N x e1 / |N|^2
        double m1[3] = {(e2[1] * N[2] - e2[2] * N[1]) / nn, (e2[2] * N[0] - e2[0] * N[2]) / nn, (e2[0] * N[1] - e2[1] * N[0]) / nn};
        double m2[3] = {(N[1] * e1[2] - N[2] * e1[1]) / nn, (N[2] * e1[0] - N[0] * e1[2]) / nn, (N[0] * e1[1] - N[1] * e1[0]) / nn};
        const double v0[3] = {v[0][0], v[0][1], v[0][2]};
        r[0] = (float)n[0]; r[1] = (float)n[1]; r[2] = (float)n[2];
        r[3] = (float)(n[0] * v0[0] + n[1] * v0[1] + n[2] * v0[2]);
        r[4] = (float)m1[0]; r[5] = (float)m1[1]; r[6] = (float)m1[2];
        r[7] = (float)-(m1[0] * v0[0] + m1[1] * v0[1] + m1[2] * v0[2]);
        r[8] = (float)m2[0]; r[9] = (float)m2[1]; r[10] = (float)m2[2];
        r[11] = (float)-(m2[0] * v0[0] + m2[1] * v0[1] + m2[2] * v0[2]);
    }
    float* b = out.bounds.data() + 6 * slot;
    for (int k = 0; k < 3; ++k) { b[k] = std::min(v[0][k], std::min(v[1][k], v[2][k])); b[3 + k] = std::max(v[0][k], std::max(v[1][k], v[2][k])); }
    out.obj[slot] = (int32_t)oi;
}
}  // namespace

void build_tri_records(const std::vector<rt_object>& objects, const std::vector<HostMesh>& meshes, TriRecords& out) {
    out = TriRecords();
    // pass 1: which triangles are kept (all three indices valid) and where they go; pass 2 fills the slots, large meshes on
    // several threads (every triangle is independent and writes only its own slot: same bytes as the sequential loop)
    struct Job { size_t oi, t, slot; };
    std::vector<Job> jobs;
    for (size_t oi = 0; oi < objects.size() && oi < meshes.size(); ++oi) {
        const rt_object& o = objects[oi];
        const HostMesh& m = meshes[oi];
        if (o.type != RT_OBJ_MESH || m.empty()) continue;
        const int nv = (int)(m.vertices.size() / 3);
        jobs.reserve(jobs.size() + m.indices.size() / 3);
        for (size_t t = 0; t + 2 < m.indices.size(); t += 3) {
            const int ia = m.indices[t], ib = m.indices[t + 1], ic = m.indices[t + 2];
            if (ia < 0 || ib < 0 || ic < 0 || ia >= nv || ib >= nv || ic >= nv) continue;
            jobs.push_back({oi, t, jobs.size()});
        }
    }
    const size_t n = jobs.size();
    out.rec.resize(12 * n); out.bounds.resize(6 * n); out.obj.resize(n);
    auto run = [&](size_t a, size_t b) {
        for (size_t j = a; j < b; ++j) tri_record(objects[jobs[j].oi], meshes[jobs[j].oi], jobs[j].t, jobs[j].oi, jobs[j].slot, out);
    };
    size_t threads = n >= ((size_t)1 << 16) ? std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 32) : 1;
    std::vector<std::thread> pool;
    const size_t chunk = (n + threads - 1) / (threads ? threads : 1);
    for (size_t k = 1; k < threads; ++k) {
        const size_t a = std::min(n, k * chunk), b = std::min(n, (k + 1) * chunk);
        if (a >= b) break;
        try { pool.emplace_back(run, a, b); } catch (const std::system_error&) { run(a, b); }   // no thread to be had: do it here
    }
    run(0, std::min(n, chunk));
    for (std::thread& th : pool) th.join();
}

bool load_obj(const std::string& path, HostMesh& out, std::string& err) {
    out = HostMesh();
    std::ifstream f(path, std::ios::binary);
    if (!f.good()) { err = "cannot open mesh file: " + path; return false; }
    out.file = path;
    std::string line;
    size_t lineno = 0;
    while (std::getline(f, line)) {
        ++lineno;
        const char* p = line.c_str();
        while (*p == ' ' || *p == '\t') ++p;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            char* e = nullptr;
            const char* q = p + 1;
            for (int k = 0; k < 3; ++k) {
                const double x = strtod(q, &e);
                if (e == q) { err = path + ":" + std::to_string(lineno) + ": bad vertex"; return false; }
                out.vertices.push_back((float)x);
                q = e;
            }
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            std::vector<int32_t> poly;
            const char* q = p + 1;
            for (;;) {
                while (*q == ' ' || *q == '\t' || *q == '\r') ++q;
                if (!*q) break;
                char* e = nullptr;
                long i = strtol(q, &e, 10);
                if (e == q) { err = path + ":" + std::to_string(lineno) + ": bad face index"; return false; }
                const long nv = (long)(out.vertices.size() / 3);
                if (i < 0) i = nv + i; else i = i - 1;                    // OBJ is 1-based; negative = relative
                if (i < 0 || i >= nv) { err = path + ":" + std::to_string(lineno) + ": face index out of range"; return false; }
                poly.push_back((int32_t)i);
                q = e;
                while (*q && *q != ' ' && *q != '\t' && *q != '\r') ++q;  // skip /vt/vn
            }
            for (size_t k = 1; k + 1 < poly.size(); ++k) { out.indices.push_back(poly[0]); out.indices.push_back(poly[k]); out.indices.push_back(poly[k + 1]); }
        }
    }
    return true;
}

bool save_obj(const std::string& path, const HostMesh& mesh, std::string& err) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write mesh file: " + path; return false; }
    for (size_t i = 0; i + 2 < mesh.vertices.size(); i += 3)
        fprintf(f, "v %.9g %.9g %.9g\n", mesh.vertices[i], mesh.vertices[i + 1], mesh.vertices[i + 2]);
    for (size_t i = 0; i + 2 < mesh.indices.size(); i += 3)
        fprintf(f, "f %d %d %d\n", mesh.indices[i] + 1, mesh.indices[i + 1] + 1, mesh.indices[i + 2] + 1);
    fclose(f);
    return true;
}

}  // namespace rtb
