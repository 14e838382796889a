// TEST STAND-IN for the reference's src/voxel_loc.hpp: the types of its lines 63-127 that immesh_shim.hpp uses (same member names).
#pragma once
#include <cstdint>
#include <functional>
#include <Eigen/Dense>
#include <pcl/point_types.h>
typedef struct ptpl {
    Eigen::Vector3d point, normal, center;
    Eigen::Matrix<double, 6, 6> plane_var;
    int layer;
    double d, eigen_value;
    bool is_valid;
} ptpl;
typedef struct Point_with_var {
    Eigen::Vector3d m_point, m_point_world;
    Eigen::Matrix3d m_var;
} Point_with_var;
class VOXEL_LOC {
  public:
    int64_t x, y, z;
    VOXEL_LOC(int64_t vx = 0, int64_t vy = 0, int64_t vz = 0) : x(vx), y(vy), z(vz) {}
    bool operator==(const VOXEL_LOC& o) const { return x == o.x && y == o.y && z == o.z; }
};
namespace std {
template <> struct hash<VOXEL_LOC> {
    size_t operator()(const VOXEL_LOC& s) const { return (size_t)((((s.z) * 116101) % 10000000000LL + (s.y)) * 116101 % 10000000000LL + (s.x)); }
};
}  // namespace std
class OctoTree {};
