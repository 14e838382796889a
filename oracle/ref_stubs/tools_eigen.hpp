// ORACLE / TEST INFRASTRUCTURE.  Stand-in for the reference's src/tools/tools_eigen.hpp (which pulls in <Eigen/Eigen>,
// not installed here).  ikd_Tree.h uses exactly one Eigen name: Eigen::aligned_allocator in the PointVector alias
// (ikd_Tree.h:108).  An allocator does not influence search results, so std::allocator is an exact substitute.
#pragma once
#include <memory>
#include <vector>
namespace Eigen {
template <class T> using aligned_allocator = std::allocator<T>;
}
