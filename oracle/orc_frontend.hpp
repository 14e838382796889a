// ORACLE (test infrastructure, NOT product code): CPU restatement of the step in FRONT of the hot path:
// the pcl::VoxelGrid centroid down-sampling of the undistorted scan,
//     m_downSizeFilterSurf.setLeafSize(l, l, l); .setInputCloud(m_feats_undistort); .filter(*m_feats_down_body);
//     (/root/reference/src/voxel_mapping.cpp:1715, :1888-1889; the mesher's copy: src/ImMesh_mesh_reconstruction.cpp:335-338)
// PCL is a third-party dependency ABSENT from /root/reference and unpinned (CMakeLists.txt:57 "find_package(PCL 1.6 ...)"),
// so this restates the published algorithm of pcl::VoxelGrid<PointT>::applyFilter (filters/impl/voxel_grid.hpp, PCL 1.8-1.12,
// default parameters: downsample_all_data_ = true, min_points_per_voxel_ = 0, filter_field_name_ empty):
//   1. inverse_leaf = 1.0f / leaf                                   (float, Eigen::Array4f)
//   2. min_p / max_p = component-wise min / max of the finite points   (getMinMax3D)
//   3. d_x = (int64)((max_p.x - min_p.x) * inverse_leaf) + 1 (...); if d_x*d_y*d_z > INT_MAX: warn, output = input
//   4. min_b = (int)floor(min_p * inverse_leaf), max_b likewise; div_b = max_b - min_b + 1; divb_mul = (1, div_b.x, div_b.x*div_b.y)
//   5. per point: ijk = (int)(floor(p * inverse_leaf) - (float)min_b);  idx = ijk . divb_mul
//   6. sort (idx, point index) by idx; one output point per run of equal idx: CentroidPoint -> float sums of x, y, z in run order,
//      divided by (float)count; output runs in ascending idx.
// PARITY UNPINNED, with one DEFINED choice: PCL sorts with std::sort on idx alone, so the order of the points inside a run (and
// with it the last ulp of a float centroid) is whatever introsort leaves; here, and in the CUDA path, the order inside a run is
// ascending point index (stable sort).
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <vector>

#include "orc_math.hpp"

namespace orc {

// KITTI laser calibration, voxel_mapping.cpp:1844-1859 (preprocess/calib_laser): in place on packed float xyz.
//   range = sqrt(x*x + y*y + z*z)   float products and sum, std::sqrt(float) (the file is compiled under `using namespace std`,
//                                   include/common_lib.h:23, so the float overloads are selected), then widened to double
//   vertical_angle = asin(z / range) + deg2rad(0.15)        double; deg2rad(x) = x * PI_M / 180.0, PI_M = 3.14159265358 (common_lib.h:34)
//   horizon_angle  = atan2(y, x)                            float arguments -> float result, widened
//   z = range * sin(vertical_angle); project_len = range * cos(vertical_angle); x = project_len * cos(horizon_angle); y = ... sin(...)
// sin / cos / asin / atan2 are the arithmetic-only implementations of orc_math.hpp (libm results differ in the last ulp between
// glibc and CUDA); tests/test_oracle_crosscheck.py compares with numpy / glibc.  PARITY UNPINNED (no reference fixture exists).
inline void kitti_calib(float* pts, int n) {
    for (int i = 0; i < n; ++i) {
        float* p = pts + 3 * (size_t)i;
        const float fx = p[0], fy = p[1], fz = p[2];
        const double range = (double)std::sqrt((fx * fx + fy * fy) + fz * fz);
        const double calib_vertical_angle = 0.15 * 3.14159265358 / 180.0;
        const double vertical_angle = det_asin((double)fz / range) + calib_vertical_angle;
        const double horizon_angle = (double)(float)det_atan2((double)fy, (double)fx);
        double sv, cv, sh, ch;
        det_sincos(vertical_angle, &sv, &cv);
        det_sincos(horizon_angle, &sh, &ch);
        p[2] = (float)(range * sv);
        const double project_len = range * cv;
        p[0] = (float)(project_len * ch);
        p[1] = (float)(project_len * sh);
    }
}

struct VoxelGridResult {
    std::vector<float> out;   // [m][3]
    int leaf_too_small = 0;   // PCL's "Leaf size is too small for the input dataset" branch: output = input
    int min_b[3] = {0, 0, 0}, div_b[3] = {0, 0, 0};
};

inline VoxelGridResult voxel_grid_filter(const float* pts, int n, float leaf) {
    VoxelGridResult R;
    const float inv = 1.0f / leaf;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int n_fin = 0;
    for (int i = 0; i < n; ++i) {
        const float* p = pts + 3 * (size_t)i;
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], p[a]); mx[a] = std::max(mx[a], p[a]); }
        ++n_fin;
    }
    if (n_fin == 0) return R;
    const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > (int64_t)INT_MAX) {
        R.leaf_too_small = 1;
        R.out.assign(pts, pts + 3 * (size_t)n);
        return R;
    }
    int max_b[3];
    for (int a = 0; a < 3; ++a) {
        R.min_b[a] = (int)std::floor(mn[a] * inv);
        max_b[a] = (int)std::floor(mx[a] * inv);
        R.div_b[a] = max_b[a] - R.min_b[a] + 1;
    }
    const int mul[3] = {1, R.div_b[0], R.div_b[0] * R.div_b[1]};
    std::vector<std::pair<unsigned, unsigned>> iv;   // (idx, point index)
    iv.reserve(n_fin);
    for (int i = 0; i < n; ++i) {
        const float* p = pts + 3 * (size_t)i;
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        int idx = 0;
        for (int a = 0; a < 3; ++a) {
            const int ijk = (int)(std::floor(p[a] * inv) - (float)R.min_b[a]);
            idx += ijk * mul[a];
        }
        iv.emplace_back((unsigned)idx, (unsigned)i);
    }
    std::stable_sort(iv.begin(), iv.end(), [](const std::pair<unsigned, unsigned>& a, const std::pair<unsigned, unsigned>& b) { return a.first < b.first; });
    size_t s = 0;
    while (s < iv.size()) {
        size_t e = s;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        while (e < iv.size() && iv[e].first == iv[s].first) {
            const float* p = pts + 3 * (size_t)iv[e].second;
            sx += p[0]; sy += p[1]; sz += p[2];
            ++e;
        }
        const float cnt = (float)(e - s);
        R.out.push_back(sx / cnt);
        R.out.push_back(sy / cnt);
        R.out.push_back(sz / cnt);
        s = e;
    }
    return R;
}

}  // namespace orc
