// ORACLE / TEST INFRASTRUCTURE -- thin C entry points over the REFERENCE's own ikd-Tree (include/ikd-Tree/ikd_Tree.{h,cpp},
// compiled unmodified from /root/reference by oracle/Makefile.ref into oracle/_ref/libref_ikd.so).  Used by tests/ and
// tools/make_golden.py to pin the oracle's kNN restatement (orc_mesh.hpp: MeshOracle::knn) and the CUDA kNN against the
// real reference code.  Mirrors the mesher's usage: KD_TREE<ikdTree_PointType>, Add_Point(pt, false) one vertex at a time
// (pointcloud_rgbd.cpp:540), Nearest_Search(pt, k, pts, d2[, max_dist]) (pointcloud_rgbd.cpp:509, mesh_rec_geometry.cpp:350).
#include "ikd_Tree.h"
#include <cstdint>

using Tree = KD_TREE<ikdTree_PointType>;

extern "C" {
void* ref_ikd_create() { return new Tree(); }   // default ctor parameters, as Global_map's member (pointcloud_rgbd.hpp:249)
void ref_ikd_destroy(void* h) { delete static_cast<Tree*>(h); }

// one Add_Point per vertex, ids first_id, first_id+1, ... (m_pt_idx = m_rgb_pts_vec.size(), pointcloud_rgbd.cpp:523-540)
void ref_ikd_add(void* h, const float* xyz, int n, int64_t first_id) {
    Tree* t = static_cast<Tree*>(h);
    for (int i = 0; i < n; i++) {
        ikdTree_PointType p(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        p.m_pt_idx = first_id + i;
        t->Add_Point(p, false);
    }
}

// idx[nq][k] (-1 padded), d2[nq][k] (inf padded), cnt[nq]
void ref_ikd_knn(void* h, const float* q, int nq, int k, double max_dist, int64_t* idx, float* d2, int* cnt) {
    Tree* t = static_cast<Tree*>(h);
    Tree::PointVector pts;
    std::vector<float> dist;
    for (int i = 0; i < nq; i++) {
        ikdTree_PointType p(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
        pts.clear();
        dist.clear();
        if (t->Root_Node != nullptr) t->Nearest_Search(p, k, pts, dist, max_dist);
        cnt[i] = (int)pts.size();
        for (int j = 0; j < k; j++) {
            idx[(size_t)i * k + j] = j < (int)pts.size() ? pts[j].m_pt_idx : -1;
            d2[(size_t)i * k + j] = j < (int)pts.size() ? dist[j] : INFINITY;
        }
    }
}
int ref_ikd_size(void* h) { return static_cast<Tree*>(h)->size(); }
}
