// immesh_shim.hpp -- header-only C++ shim that gives the reference ROS node the SAME free-function signatures it
// already calls (src/voxel_mapping.hpp:80-105, src/ImMesh_mesh_reconstruction.cpp:92, include/ikd-Tree/ikd_Tree.h:306)
// and forwards them to the C ABI of libimmesh_b200.so.  It is compiled inside the reference tree (it needs the
// reference's own headers: Eigen, PCL point types, voxel_loc.hpp); INTEGRATION.md shows where it is included.  In this
// repository it is type-checked and linked against minimal stand-ins of those headers (tests/shim_stubs/, tests/test_shim_compile.py).
// Nothing here touches CUDA.
//
// Usage in the reference:  #define IMMESH_B200_SHIM before including voxel_mapping.hpp, link -limmesh_b200.
#pragma once
#ifdef IMMESH_B200_SHIM
#include <cmath>
#include <cstdint>
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

// ---- mesher seams -----------------------------------------------------------------------------------------------
// The reference keeps the mesher's parameters in globals set by main() (src/ImMesh_node.cpp:93-98, :255-257):
extern double minimum_pts;            // m_meshing_points_minimum_scale * distance_scale
extern double g_meshing_voxel_size;   // m_meshing_voxel_resolution * distance_scale
extern int appending_pts_frame;       // m_meshing_number_of_pts_append_to_map
namespace immesh_shim {
inline immesh_mesh_t*& mesh_handle() {
    static immesh_mesh_t* h = nullptr;
    return h;
}
inline immesh_mesh_t* mesh_handle_or_create() {
    if (!mesh_handle()) {
        immesh_mesh_config c{};
        c.points_minimum_scale = minimum_pts;
        c.voxel_resolution = g_meshing_voxel_size;
        c.number_of_pts_append_to_map = appending_pts_frame;
        if (immesh_mesh_create(&c, &mesh_handle()) != IMMESH_OK) mesh_handle() = nullptr;
    }
    return mesh_handle();
}
// KD_TREE<ikdTree_PointType>::Nearest_Search(point, k_nearest, Nearest_Points, Point_Distance, max_dist) (include/ikd-Tree/ikd_Tree.h:306)
// over the mesher's vertices: same outputs -- the k nearest vertices in ascending distance with their m_pt_idx, squared float distances.
template <class PointType, class PointVector>
inline void Nearest_Search(PointType point, int k_nearest, PointVector& Nearest_Points, std::vector<float>& Point_Distance, double max_dist = INFINITY) {
    Nearest_Points.clear();
    Point_Distance.clear();
    immesh_mesh_t* h = mesh_handle_or_create();
    if (!h || k_nearest < 1) return;
    const float q[3] = {point.x, point.y, point.z};
    std::vector<int32_t> idx((size_t)k_nearest);
    std::vector<float> d2((size_t)k_nearest);
    std::vector<float> pos;
    if (immesh_knn(h, q, 1, k_nearest, max_dist, idx.data(), d2.data()) != IMMESH_OK) return;
    int64_t cnt[8];
    immesh_mesh_counts(h, cnt);
    pos.resize((size_t)cnt[0] * 3);
    immesh_mesh_snapshot(h, pos.data(), nullptr, nullptr);
    for (int i = 0; i < k_nearest && idx[i] >= 0; ++i) {
        PointType p(pos[(size_t)idx[i] * 3], pos[(size_t)idx[i] * 3 + 1], pos[(size_t)idx[i] * 3 + 2]);
        p.m_pt_idx = idx[i];
        Nearest_Points.push_back(p);
        Point_Distance.push_back(d2[i]);
    }
}
}  // namespace immesh_shim

// src/ImMesh_mesh_reconstruction.cpp:92: one frame of the voxel-wise incremental mesher.  frame_pts is the full-resolution scan already in
// the world frame (transformLidar's output, float xyz + intensity); pose_q is only logged by the reference (:99-105), pose_t is the
// sensor position used by the facet orientation (correct_triangle_index).
inline void incremental_mesh_reconstruction(pcl::PointCloud<pcl::PointXYZI>::Ptr frame_pts, Eigen::Quaterniond pose_q, Eigen::Vector3d pose_t, int frame_idx) {
    (void)pose_q;
    immesh_mesh_t* h = immesh_shim::mesh_handle_or_create();
    if (!h || !frame_pts) return;
    const size_t n = frame_pts->points.size();
    std::vector<float> xyz(n * 3);
    for (size_t i = 0; i < n; ++i) {
        xyz[i * 3 + 0] = frame_pts->points[i].x; xyz[i * 3 + 1] = frame_pts->points[i].y; xyz[i * 3 + 2] = frame_pts->points[i].z;
    }
    const double t[3] = {pose_t[0], pose_t[1], pose_t[2]};
    immesh_mesh_push_frame(h, xyz.data(), (int)n, t, frame_idx);
}
#endif  // IMMESH_B200_SHIM
