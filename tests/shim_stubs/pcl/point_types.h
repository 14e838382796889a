// TEST STAND-IN (not PCL): point / cloud types with the members immesh_shim.hpp touches.
#pragma once
#include <memory>
#include <vector>
namespace pcl {
struct PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; };
template <class P> struct PointCloud {
    typedef std::shared_ptr<PointCloud<P>> Ptr;
    std::vector<P> points;
};
}  // namespace pcl
