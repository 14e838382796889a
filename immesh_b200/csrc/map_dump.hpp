// immesh_b200 -- host-side canonical dump of the VoxelMap pools (diagnostics / parity tests).
// Row format (45 doubles per octree node, roots by ascending key, nodes in pre-order):
// kx ky kz path layer init_octo is_plane update_enable n_temp new_points 0 | centre[3] normal[3] d radius min_eig
// points_size | voxel_centre[3] | plane_var_upper[21]
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "voxelmap.cuh"

namespace immesh {
inline void dump_node_rec(long long kx, long long ky, long long kz, int nd, int64_t path, const NodeRec* nodes, const PlaneRec* planes,
                          std::vector<double>& out) {
    const NodeRec& n = nodes[nd];
    const PlaneRec& p = planes[nd];
    double row[45];
    int c = 0;
    row[c++] = (double)kx; row[c++] = (double)ky; row[c++] = (double)kz;
    row[c++] = (double)path; row[c++] = n.layer; row[c++] = n.init_octo; row[c++] = p.is_plane;
    row[c++] = n.update_enable; row[c++] = n.n_pts; row[c++] = n.new_points; row[c++] = 0;
    const bool pl = p.is_plane != 0;
    for (int i = 0; i < 3; ++i) row[c++] = p.center[i];
    for (int i = 0; i < 3; ++i) row[c++] = pl ? p.normal[i] : 0.0;
    row[c++] = pl ? p.d : 0.0; row[c++] = pl ? p.radius : 0.0; row[c++] = pl ? p.min_eig : 0.0;
    row[c++] = p.points_size;
    for (int i = 0; i < 3; ++i) row[c++] = n.vc[i];
    for (int i = 0; i < 21; ++i) row[c++] = pl ? p.pv[i] : 0.0;
    out.insert(out.end(), row, row + 45);
    for (int l = 0; l < 8; ++l)
        if (n.children[l] >= 0) dump_node_rec(kx, ky, kz, n.children[l], path * 9 + (l + 1), nodes, planes, out);
}
inline int64_t dump_voxelmap(const unsigned long long* keys, const int* root_node, size_t cap, const NodeRec* nodes, const PlaneRec* planes,
                             double* rows, int64_t cap_rows) {
    struct R { long long x, y, z; int node; };
    std::vector<R> roots;
    for (size_t s = 0; s < cap; ++s)
        if (keys[s] != IM_EMPTY_KEY && root_node[s] >= 0) {
            R r;
            unpack_key(keys[s], &r.x, &r.y, &r.z);
            r.node = root_node[s];
            roots.push_back(r);
        }
    std::sort(roots.begin(), roots.end(), [](const R& a, const R& b) {
        if (a.x != b.x) return a.x < b.x;
        if (a.y != b.y) return a.y < b.y;
        return a.z < b.z;
    });
    std::vector<double> out;
    for (const R& r : roots) dump_node_rec(r.x, r.y, r.z, r.node, 0, nodes, planes, out);
    const int64_t nrows = (int64_t)(out.size() / 45);
    if (rows && nrows <= cap_rows) std::copy(out.begin(), out.end(), rows);
    return nrows;
}
}  // namespace immesh
