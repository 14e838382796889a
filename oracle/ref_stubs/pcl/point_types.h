// ORACLE / TEST INFRASTRUCTURE.  Stand-in for <pcl/point_types.h> (PCL is not installed in this image) used ONLY to
// compile the reference's own include/ikd-Tree/ikd_Tree.cpp, unmodified and from where it lies under /root/reference,
// into oracle/_ref/libref_ikd.so (see oracle/Makefile.ref).  ikd_Tree.cpp needs these three names solely for its explicit
// template instantiations (ikd_Tree.cpp:1825-1827); the mesher uses KD_TREE<ikdTree_PointType> (pointcloud_rgbd.hpp:232),
// whose point type is defined inside ikd_Tree.h itself.  Field layout follows PCL's public documentation (x,y,z float).
#pragma once
#include <cstring>   // the real header provides memcpy/memset transitively; ikd_Tree.cpp relies on that
#include <cmath>
namespace pcl {
struct PointXYZ { float x, y, z; };
struct PointXYZI { float x, y, z, intensity; };
struct PointXYZINormal { float x, y, z, intensity, normal_x, normal_y, normal_z, curvature; };
}  // namespace pcl
