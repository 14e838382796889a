// A translation unit written like the reference's call sites (voxel_mapping.cpp:1274,1365; ImMesh_mesh_reconstruction.cpp:408;
// ImMesh_node.cpp:272; pointcloud_rgbd.cpp:507) against the shim: type-checks the shim's signatures and links the C ABI.
#define IMMESH_B200_SHIM
#include "immesh_shim.hpp"
double minimum_pts = 0.1, g_meshing_voxel_size = 0.4;
int appending_pts_frame = 5000;
struct ikdTree_PointType {
    float x, y, z;
    long m_pt_idx = -1;
    ikdTree_PointType(float px = 0.f, float py = 0.f, float pz = 0.f) : x(px), y(py), z(pz) {}
};
int shim_user(int run) {
    std::unordered_map<VOXEL_LOC, OctoTree*> feat_map;
    std::vector<Point_with_var> pv(3);
    std::vector<int> layer_init{5, 5, 5, 5, 5};
    std::vector<ptpl> ptpl_list;
    std::vector<Eigen::Vector3d> non_match;
    std::vector<ikdTree_PointType> nearest;
    std::vector<float> dist;
    pcl::PointCloud<pcl::PointXYZI>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZI>());
    if (run) {   // never executed by the CPU test (needs a GPU); present so that every call is instantiated and linked
        buildVoxelMap(pv, 0.5f, 2, layer_init, 100, 0.01f, feat_map);
        updateVoxelMap(pv, 0.5f, 2, layer_init, 100, 0.01f, feat_map);
        BuildResidualListOMP(feat_map, 0.5, 3.0, 2, pv, ptpl_list, non_match);
        incremental_mesh_reconstruction(cloud, Eigen::Quaterniond(), Eigen::Vector3d(), 0);
        immesh_shim::Nearest_Search(ikdTree_PointType(1.f, 2.f, 3.f), 20, nearest, dist);
    }
    return (int)ptpl_list.size();
}
int main() { return shim_user(0); }
