// immesh_b200 -- version / error string / profiling entry points.
#include <cstdio>
#include <cstring>

#include "common_host.hpp"

namespace immesh {
Profiler& profiler() {
    static Profiler p;
    return p;
}
}  // namespace immesh

extern "C" {
// save_to_ply_file (src/meshing/mesh_rec_geometry.cpp:71-131, smooth_factor == 0 branch): binary little-endian PLY, vertices as float
// x y z, one face per live triangle with the reference's orientation rule -- m_index_flip != 0 keeps (p0, p1, p2), == 0 writes
// (p0, p2, p1) (:108-121).  Host-side only: the arrays are what immesh_mesh_snapshot returns.
int immesh_write_ply(const char* path, const float* vertices, int nv, const int32_t* triangles, const int32_t* flips, int nt) {
    if (!path || (!vertices && nv > 0) || (!triangles && nt > 0) || nv < 0 || nt < 0) return immesh::im_fail(IMMESH_E_INVALID, "bad argument");
    for (int t = 0; t < nt * 3; ++t)
        if (triangles[t] < 0 || triangles[t] >= nv) return immesh::im_fail(IMMESH_E_RANGE, "triangle refers to a vertex outside the vertex array");
    FILE* f = std::fopen(path, "wb");
    if (!f) return immesh::im_fail(IMMESH_E_INVALID, "cannot open the output file");
    std::fprintf(f, "ply\nformat binary_little_endian 1.0\ncomment immesh_b200 (layout of pcl::io::savePLYFileBinary for a PolygonMesh of PointXYZ)\n"
                    "element vertex %d\nproperty float x\nproperty float y\nproperty float z\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n", nv, nt);
    if (nv > 0) std::fwrite(vertices, sizeof(float) * 3, (size_t)nv, f);
    for (int t = 0; t < nt; ++t) {
        const unsigned char three = 3;
        const int32_t a = triangles[3 * t], b = triangles[3 * t + 1], c = triangles[3 * t + 2];
        const int32_t face[3] = {a, (flips && flips[t] != 0) ? b : c, (flips && flips[t] != 0) ? c : b};
        std::fwrite(&three, 1, 1, f);
        std::fwrite(face, sizeof(int32_t), 3, f);
    }
    const bool ok = std::fclose(f) == 0;
    return ok ? IMMESH_OK : immesh::im_fail(IMMESH_E_INVALID, "write failed");
}
const char* immesh_last_error(void) { return immesh::last_error_storage().c_str(); }
const char* immesh_version(void) { return "immesh_b200 0.1.0 (sm_100a)"; }

int immesh_profile_enable(int on) {
    immesh::profiler().enabled = on != 0;
    if (on == 2) immesh::profiler().start_timeline();   // also record every launch's start / end time
    else immesh::profiler().timeline_on = false;
    return IMMESH_OK;
}
int immesh_profile_reset(void) {
    immesh::profiler().totals.clear();
    immesh::profiler().launches = 0;
    return IMMESH_OK;
}
long long immesh_launch_count(void) { return immesh::profiler().launches; }
// "name t0_ms t1_ms\n" per launch recorded since immesh_profile_enable(2); returns the number of bytes needed
int immesh_profile_timeline(char* buf, int cap) {
    std::string s;
    for (auto& sp : immesh::profiler().timeline) {
        char line[256];
        std::snprintf(line, sizeof(line), "%s %.6f %.6f\n", sp.name, sp.t0, sp.t1);
        s += line;
    }
    if (buf && cap > 0) {
        std::strncpy(buf, s.c_str(), (size_t)cap - 1);
        buf[cap - 1] = 0;
    }
    return (int)s.size() + 1;
}
// writes "name ms launches\n" lines into buf; returns the number of bytes needed
int immesh_profile_report(char* buf, int cap) {
    std::string s;
    for (auto& kv : immesh::profiler().totals) {
        char line[256];
        std::snprintf(line, sizeof(line), "%s %.6f %lld\n", kv.first.c_str(), kv.second.first, kv.second.second);
        s += line;
    }
    if (buf && cap > 0) {
        std::strncpy(buf, s.c_str(), (size_t)cap - 1);
        buf[cap - 1] = 0;
    }
    return (int)s.size() + 1;
}
}
