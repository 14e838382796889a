// immesh_shim.hpp -- header-only C++ shim that gives the reference ROS node the SAME free-function signatures it
// already calls (src/voxel_mapping.hpp:80-105, src/ImMesh_mesh_reconstruction.cpp:92, include/ikd-Tree/ikd_Tree.h:306)
// and forwards them to the C ABI of libimmesh_b200.so.  It is compiled inside the reference tree (it needs the
// reference's own headers: Eigen, PCL point types, voxel_loc.hpp), so it cannot be built in this repository's
// sandbox; INTEGRATION.md shows where it is included.  Nothing here touches CUDA.
//
// Usage in the reference:  #define IMMESH_B200_SHIM before including voxel_mapping.hpp, link -limmesh_b200.
#pragma once
#ifdef IMMESH_B200_SHIM
#include <unordered_map>
#include <vector>

#include "immesh_b200.h"
#include "voxel_loc.hpp"  // reference header: Point_with_var, ptpl, VOXEL_LOC, OctoTree

namespace immesh_shim {

// one device-resident VoxelMap per reference feat_map object (the reference owns exactly one: Voxel_mapping::m_feat_map)
inline std::unordered_map<const void*, immesh_lio_t*>& registry() {
    static std::unordered_map<const void*, immesh_lio_t*> r;
    return r;
}
inline immesh_lio_t* handle_for(const void* feat_map, float voxel_size, int max_layer, const std::vector<int>& layer_init_num, int max_points_size,
                                float planer_threshold) {
    auto it = registry().find(feat_map);
    if (it != registry().end()) return it->second;
    immesh_lio_config c{};
    c.voxel_size = voxel_size;
    c.max_layer = max_layer;
    for (int i = 0; i < 5; ++i) c.layer_init_size[i] = i < (int)layer_init_num.size() ? layer_init_num[i] : 5;
    c.max_points_size = max_points_size;
    c.min_eigen_value = planer_threshold;
    c.dept_err = 0.02; c.beam_err = 0.05;           // only used by the fused immesh_lio_* calls, not by the *_pv calls below
    c.ext_R[0] = c.ext_R[4] = c.ext_R[8] = 1.0;
    c.max_iteration = 4;
    immesh_lio_t* h = nullptr;
    if (immesh_lio_create(&c, &h) != IMMESH_OK) return nullptr;
    registry()[feat_map] = h;
    return h;
}
inline void flatten(const std::vector<Point_with_var>& in, bool world_field, std::vector<double>& pts, std::vector<double>& var) {
    pts.resize(in.size() * 3);
    var.resize(in.size() * 9);
    for (size_t i = 0; i < in.size(); ++i) {
        const Eigen::Vector3d& p = world_field ? in[i].m_point_world : in[i].m_point;
        for (int j = 0; j < 3; ++j) pts[i * 3 + j] = p[j];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) var[i * 9 + r * 3 + c] = in[i].m_var(r, c);
    }
}
}  // namespace immesh_shim

// src/voxel_mapping.hpp:80-82
inline void buildVoxelMap(const std::vector<Point_with_var>& input_points, const float voxel_size, const int max_layer,
                          const std::vector<int>& layer_init_num, const int max_points_size, const float planer_threshold,
                          std::unordered_map<VOXEL_LOC, OctoTree*>& feat_map) {
    immesh_lio_t* h = immesh_shim::handle_for(&feat_map, voxel_size, max_layer, layer_init_num, max_points_size, planer_threshold);
    std::vector<double> pts, var;
    immesh_shim::flatten(input_points, false, pts, var);
    if (h) immesh_voxelmap_build_pv(h, pts.data(), var.data(), (int)input_points.size());
}
// src/voxel_mapping.hpp:90-92
inline void updateVoxelMap(const std::vector<Point_with_var>& input_points, const float voxel_size, const int max_layer,
                           const std::vector<int>& layer_init_num, const int max_points_size, const float planer_threshold,
                           std::unordered_map<VOXEL_LOC, OctoTree*>& feat_map) {
    immesh_lio_t* h = immesh_shim::handle_for(&feat_map, voxel_size, max_layer, layer_init_num, max_points_size, planer_threshold);
    std::vector<double> pts, var;
    immesh_shim::flatten(input_points, false, pts, var);
    if (h) immesh_voxelmap_update_pv(h, pts.data(), var.data(), (int)input_points.size());
}
// src/voxel_mapping.hpp:103-105 (non_match is never filled by the reference either)
inline void BuildResidualListOMP(const std::unordered_map<VOXEL_LOC, OctoTree*>& voxel_map, const double voxel_size, const double sigma_num,
                                 const int max_layer, const std::vector<Point_with_var>& pv_list, std::vector<ptpl>& ptpl_list,
                                 std::vector<Eigen::Vector3d>& non_match) {
    (void)voxel_size; (void)sigma_num; (void)max_layer; (void)non_match;
    ptpl_list.clear();
    auto it = immesh_shim::registry().find(&voxel_map);
    if (it == immesh_shim::registry().end()) return;
    const int n = (int)pv_list.size();
    std::vector<double> body(n * 3), world, var;
    immesh_shim::flatten(pv_list, true, world, var);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j) body[i * 3 + j] = pv_list[i].m_point[j];
    std::vector<int> il(2 * (size_t)n);
    std::vector<double> vals(31 * (size_t)n);
    int m = 0;
    if (immesh_residual_build_pv(it->second, body.data(), world.data(), var.data(), n, il.data(), vals.data(), n, &m) != IMMESH_OK) return;
    ptpl_list.resize(m);
    for (int k = 0; k < m; ++k) {
        const double* v = &vals[31 * (size_t)k];
        ptpl& o = ptpl_list[k];
        o.point << v[0], v[1], v[2];
        o.normal << v[3], v[4], v[5];
        o.center << v[6], v[7], v[8];
        o.d = v[9];
        o.layer = il[2 * k + 1];
        int e = 10;
        for (int r = 0; r < 6; ++r)
            for (int c = r; c < 6; ++c, ++e) { o.plane_var(r, c) = v[e]; o.plane_var(c, r) = v[e]; }
    }
}
#endif  // IMMESH_B200_SHIM
