// immesh_b200 -- device-resident voxel-wise incremental mesher.
//
// Reference semantics being reproduced (all /root/reference):
//   append_points_to_global_map      src/meshing/r3live/pointcloud_rgbd.cpp:411-552
//   KD_TREE::Nearest_Search (exact kNN, float)   include/ikd-Tree/ikd_Tree.cpp:441-476, 1097-1279, 1722-1727
//   retrieve_pts_in_voxels / retrieve_neighbor_pts_kdtree   pointcloud_rgbd.cpp:397-407, mesh_rec_geometry.cpp:336-377
//   delaunay_triangulation + is_face_is_ok   src/meshing/mesh_rec_geometry.cpp:24-57, 174-295
//   find_relative_triangulation_combination (pull)   src/meshing/r3live/triangle.hpp:223-246
//   triangle_compare (commit), correct_triangle_index   mesh_rec_geometry.cpp:137-172, 399-433
//   remove_triangle_list / insert_triangle (push)    triangle.hpp:164-221, 330-395
//   incremental_mesh_reconstruction (driver)         src/ImMesh_mesh_reconstruction.cpp:92-267
//
// Data layout in HBM: vertex SoA (float4 position, double3 smoothed position, intrusive per-voxel list,
// incidence-list head), two open-addressed hashes (xi-cell -> vertex, mesh voxel -> voxel record; the slot
// index IS the voxel id), triangle pool (sorted id triple + alive flag, 3 intrusive incidence links, flip
// word) with an open-addressed triple hash.  Per frame: candidate arrays, activated / work voxel lists,
// per-voxel dilated id lists, add / remove lists.
//
// Parallel decomposition: vertex append = lexicographically-first maximal independent set in scan order,
// resolved by priority polling (identical to the sequential greedy loop of the reference); dilation / kNN,
// PCA + Delaunay (Bowyer-Watson with block-wide conflict search and exact integer predicates) and
// pull/commit = one thread block per activated voxel; push = one thread per list entry.
#pragma once
#include <cuda_runtime.h>  // vector types (float4, int4, int2); host-only builds use the same headers
#include <cstring>

#include "hd_math.cuh"
#include "voxelmap.cuh"  // atomics + key packing helpers

#if defined(__CUDA_ARCH__)
#define IM_SYNCBLOCK_M() __syncthreads()
#else
#define IM_SYNCBLOCK_M()
#endif

namespace immesh {

struct MeshParams {
    double xi;         // m_minimum_pts_size
    double res;        // m_voxel_resolution
    double accept;     // g_kd_tree_accept_pt_dis = 1.25 * res (mesh_rec_geometry.cpp:343)
    double knn_max;    // 2 * accept * (1 + 1e-6): beyond this both filters of the dilation ignore a neighbour
    double inv_q;      // 2^22 / res: snapping factor of the exact Delaunay predicates
    int append_target; // appending_pts_frame
};

enum : int { CAND_UNDECIDED = 0, CAND_ACCEPT = 1, CAND_REJECT = 2 };
#define IM_MAXD 1024        // max dilated vertices per voxel
#define IM_MAXG 2048        // max gathered kNN candidates per voxel neighbourhood
#define IM_MAXT (3 * IM_MAXD + 8)
#define IM_MAXPULL 4096
#define IM_MAXF 512         // facets of a voxel with <= 256 dilated vertices (2n - 5 at most)
#define IM_MAXIN 256        // max in-voxel vertices
#define IM_CONF_K 24        // stored earlier-conflict candidates per candidate
#define IM_VCHUNKS 16        // 16 chunks x 16 vertices = IM_MAXIN vertices per mesh voxel

enum : int {
    IM_MERR_VERT_POOL = 1,
    IM_MERR_TRI_POOL = 2,
    IM_MERR_HASH_FULL = 4,
    IM_MERR_VOXEL_CAP = 8,   // a per-voxel working set exceeded IM_MAXD / IM_MAXG / IM_MAXPULL / IM_MAXIN
    IM_MERR_LIST_CAP = 16,   // add/remove/work list overflow
    IM_MERR_KEY_RANGE = 32,
    IM_MERR_PRIO_RANGE = 64, // activated voxel farther than 1024 voxels from the sensor (flip priority)
    IM_MERR_PEER_TIMEOUT = 128,   // sharded mode: a peer rank's epoch flag did not arrive (peer_win.cuh)
};

struct MeshDev {
    // vertices
    float4* vpos;        // xyz, w unused
    double* vsmooth;     // [max_v][3]
    int* v_tri_head;     // head of the incidence list
    int max_v;
    // xi-grid hash: cell -> vertex id
    unsigned long long* gkeys;
    int* gval;
    unsigned int gmask;
    // mesh-voxel hash (slot = voxel id)
    unsigned long long* vkeys;
    unsigned int vmask;
    int* vox_chunk;      // [cap][IM_VCHUNKS] ids of the voxel's vertex chunks (-1 = none); vertex k of a voxel lives in
                         // vchunk_pts[vox_chunk[k / 16] * 16 + k % 16] -> a voxel's vertices are fetched with a few wide, independent loads
    float4* vchunk_pts;  // [max_vchunks][16] xyz + vertex id (bit-cast in w)
    int max_vchunks;
    int* vox_count;
    int* vox_meshing_times;
    int* vox_new_added;
    int* vox_frame;
    double* vox_short_axis;  // [cap][3]
    // triangles
    int4* tri;            // a<b<c, alive
    int* tri_next;        // [max_t][3]
    unsigned long long* tri_flip;
    int* thash;           // open addressing on the triple, value = triangle index
    unsigned int tmask;
    int max_t;
    // counters: 0 n_vertices, 1 n_tris_alloc, 2 n_live, 3 err, 4 n_voxels, 5 n_act, 6 n_work, 7 n_add, 8 n_remove,
    //           9 n_undecided, 10 frame_new_vertices, 11..13 key bbox min, 14..16 key bbox max, 17 bbox valid
    int* cnt;
};

struct FramePose {        // per-frame sensor position (device resident so that a frame can be queued behind the localization)
    double pose_t[3];
    long long prio_origin[3];  // floor(pose_t / res) - 1024: origin of the 11-bit-per-axis voxel rank used by the flip priority
};
struct XSmooth { int id, pad; double x, y, z; };   // 32 B
// Per-frame inputs of the launch sequence, device resident (one small H2D per frame): the captured CUDA graph of a frame is
// replayed without touching its nodes, every kernel takes the frame's values from here (frame_load_dyn).
struct FrameDyn {
    const float* pts;     // world-frame scan of this frame (staging slot, or the caller's device buffer)
    const float* body;    // body-frame scan when the frame is handed over by the localization handle (transformed into pts first)
    FramePose* fp;
    int n, step, m, frame;
    unsigned int cmask;
    int pose_idx;         // slot of LioCtrl::pose_ring holding the pose the scan converged to
    unsigned long long epoch;   // peer-window epoch of this frame (sharded mode)
};
struct FrameBuf {
    const float* pts;     // [n][3] world-frame scan
    int n, step, m;       // m = number of candidates = ceil(n / step)
    int frame;            // internal monotonically increasing frame counter
    const FramePose* fp;
    const FrameDyn* dyn;  // device: the seven per-frame fields above + cmask + epoch are taken from here by the kernels
    unsigned long long epoch;
    // candidates
    unsigned long long* cand_gkey;
    int* cand_vslot;
    int* cand_status;
    int* cand_scan;       // exclusive scan of accept flags
    int* cand_conf;       // [m][IM_CONF_K]
    int* cand_nconf;      // count, or -1 when the list overflowed
    int* cand_next;       // next candidate in the same xi-cell (per-frame candidate grid)
    int* cand_pos;        // position of an accepted candidate inside its mesh voxel
    unsigned long long* ckeys;  // per-frame candidate grid
    int* chead;
    unsigned int cmask;
    int* scan_block;      // block sums for the scan
    // voxel lists
    int* act;             // activated voxel slots
    int* work;            // voxels to (re)mesh this frame
    int* work_n_ids;      // dilated set sizes
    int* work_ids;        // [max_work][IM_MAXD] ascending vertex ids
    unsigned int* work_bits;  // [max_work][IM_MAXG/32] dilation-member bitmap shared by the groups of a voxel
    int* work_ring;       // largest ring any group of the voxel gathered
    int* work_done;       // groups of the voxel that have finished
    int* ditem;           // dilation items: work slot << 5 | query group
    int max_ditem;
    int* work_nfaces;     // facets produced by the fused dilate+triangulate stage (-1: left to the large variant)
    int4* all_faces;      // compact list of this frame's new facets: (a, b, c, work slot)   [max_list]
    int* all_vref;        // compact list of (work slot << 10 | index into its dilated id list) [max_vref]
    int* pulled;          // [max_list][2] (work slot, triangle) pairs found by the pull stage
    int* fset;            // open-addressed set over all_faces (value = facet index), cleared per frame
    unsigned int fset_mask;
    int max_vref;
    double* work_axes;    // [max_work][9] short / mid / long axis of the voxel's PCA frame
    // push lists
    int* add_tri;         // [max_list][3]
    unsigned long long* add_flip;
    int* rem_tri;         // [max_list] triangle indices
    int max_cand, max_act, max_work, max_list;
    // multi-GPU: the per-voxel stage of a frame is sharded by mesh-voxel owner; what a rank produces is recorded in these
    // exchange lists (all-gathered, then applied by every rank to its replica of the store) instead of being applied
    int shard_rank, shard_n;
    int x_cap;            // capacity of each exchange list (entries)
    XSmooth* x_smooth;    // smoothed positions written by this rank's dilations      (count: cnt[32])
    int4* x_face;         // new facets (a, b, c, -) of this rank's voxels             (count: cnt[30])
    unsigned long long* x_word;   // their flip-priority words
    int4* x_rem;          // triangles (a, b, c, -) this rank's voxels want removed    (count: cnt[31])
};

IM_HD FrameBuf frame_load_dyn(const FrameBuf& F) {
    FrameBuf o = F;
    if (F.dyn) {
        const FrameDyn d = *F.dyn;
        o.pts = d.pts; o.n = d.n; o.step = d.step; o.m = d.m; o.frame = d.frame; o.fp = d.fp; o.cmask = d.cmask; o.epoch = d.epoch;
    }
    return o;
}

// ------------------------------------------------------------------ keys
IM_HD int round_key(float x, double cell) { return (int)round((double)x / cell); }  // std::round, pointcloud_rgbd.cpp:467-472
IM_HD unsigned long long pack_ikey(int x, int y, int z) {
    return ((unsigned long long)(unsigned int)(x + 1048576) << 42) | ((unsigned long long)(unsigned int)(y + 1048576) << 21) | (unsigned long long)(unsigned int)(z + 1048576);
}
IM_HD bool ikey_ok(int x, int y, int z) { return x > -1048000 && x < 1048000 && y > -1048000 && y < 1048000 && z > -1048000 && z < 1048000; }
IM_HD void unpack_ikey(unsigned long long k, int* x, int* y, int* z) {
    *x = (int)((k >> 42) & 0x1FFFFF) - 1048576;
    *y = (int)((k >> 21) & 0x1FFFFF) - 1048576;
    *z = (int)(k & 0x1FFFFF) - 1048576;
}
IM_HD int table_find(const unsigned long long* keys, unsigned int mask, unsigned long long key) {
    unsigned int s = hash_key(key) & mask;
    for (unsigned int probe = 0; probe <= mask; ++probe) {
        const unsigned long long k = keys[s];
        if (k == key) return (int)s;
        if (k == IM_EMPTY_KEY) return -1;
        s = (s + 1) & mask;
    }
    return -1;
}
IM_HD int table_insert(unsigned long long* keys, unsigned int mask, unsigned long long key, int* created) {
    unsigned int s = hash_key(key) & mask;
    *created = 0;
    for (unsigned int probe = 0; probe <= mask; ++probe) {
        unsigned long long k = keys[s];
        if (k == key) return (int)s;
        if (k == IM_EMPTY_KEY) {
            k = im_atomic_cas64(&keys[s], IM_EMPTY_KEY, key);
            if (k == IM_EMPTY_KEY) { *created = 1; return (int)s; }
            if (k == key) return (int)s;
        }
        s = (s + 1) & mask;
    }
    return -1;
}
IM_HD int im_atomic_exch(int* p, int v) {
#if defined(__CUDA_ARCH__)
    return atomicExch(p, v);
#else
    const int o = *p; *p = v; return o;
#endif
}
IM_HD int im_atomic_cas32(int* p, int cmp, int v) {
#if defined(__CUDA_ARCH__)
    return atomicCAS(p, cmp, v);
#else
    const int o = *p; if (o == cmp) *p = v; return o;
#endif
}
IM_HD int im_atomic_min(int* p, int v) {
#if defined(__CUDA_ARCH__)
    return atomicMin(p, v);
#else
    const int o = *p; if (v < o) *p = v; return o;
#endif
}
IM_HD int im_atomic_max(int* p, int v) {
#if defined(__CUDA_ARCH__)
    return atomicMax(p, v);
#else
    const int o = *p; if (v > o) *p = v; return o;
#endif
}
IM_HD unsigned long long im_atomic_max64(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return atomicMax(p, v);
#else
    const unsigned long long o = *p; if (v > o) *p = v; return o;
#endif
}
IM_HD int im_vload(const int* p) {
#if defined(__CUDA_ARCH__)
    return *(const volatile int*)p;
#else
    return *p;
#endif
}
IM_HD void im_fence() {
#if defined(__CUDA_ARCH__)
    __threadfence();
#endif
}
// float squared distance exactly as KD_TREE::calc_dist (ikd_Tree.cpp:1722-1727): ((dx*dx + dy*dy) + dz*dz), no FMA
IM_HD float dist2f(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    return (dx * dx + dy * dy) + dz * dz;
}

// ------------------------------------------------------------------ M1: candidates (thread per candidate)
// pointcloud_rgbd.cpp:464-517: grid / voxel keys, voxel get-or-create + activation, rejection against the
// vertices of earlier frames (occupied xi-cell, or an existing vertex closer than xi).
IM_HDN inline void cand_init(const MeshDev& M, const MeshParams& P, const FrameBuf& F, int c) {
    IM_STAMP(53, 0);
    const float* p = F.pts + (size_t)c * F.step * 3;
    const float px = p[0], py = p[1], pz = p[2];
    IM_STAMP(54, __float_as_int(px + py + pz));
    const int gx = round_key(px, P.xi), gy = round_key(py, P.xi), gz = round_key(pz, P.xi);
    const int bx = round_key(px, P.res), by = round_key(py, P.res), bz = round_key(pz, P.res);
    F.cand_status[c] = CAND_REJECT;
    F.cand_vslot[c] = -1;
    F.cand_nconf[c] = 0;
    F.cand_next[c] = -1;
    if (!ikey_ok(gx, gy, gz) || !ikey_ok(bx, by, bz)) { im_atomic_or(&M.cnt[3], IM_MERR_KEY_RANGE); return; }
    const unsigned long long gkey = pack_ikey(gx, gy, gz);
    F.cand_gkey[c] = gkey;
    // voxel get-or-create (:482-495); every examined point activates its voxel, accepted or not
    int created = 0;
    const int vs = table_insert(M.vkeys, M.vmask, pack_ikey(bx, by, bz), &created);
    if (vs < 0) { im_atomic_or(&M.cnt[3], IM_MERR_HASH_FULL); return; }
    IM_STAMP(55, vs);
    if (created) {
        im_atomic_add(&M.cnt[4], 1);
        im_atomic_min(&M.cnt[11], bx); im_atomic_min(&M.cnt[12], by); im_atomic_min(&M.cnt[13], bz);
        im_atomic_max(&M.cnt[14], bx); im_atomic_max(&M.cnt[15], by); im_atomic_max(&M.cnt[16], bz);
    }
    F.cand_vslot[c] = vs;
    if (im_atomic_exch(&M.vox_frame[vs], F.frame) != F.frame) {
        const int a = im_atomic_add(&M.cnt[5], 1);
        if (a < F.max_act) F.act[a] = vs; else im_atomic_or(&M.cnt[3], IM_MERR_LIST_CAP);
    }
    IM_STAMP(56, 0);
    // occupied xi-cell (:473-481)
    const int occupied = table_find(M.gkeys, M.gmask, gkey);
    IM_STAMP(57, occupied);
    if (occupied >= 0) return;
    // nearest existing vertex closer than xi (:507-517): any vertex with sqrtf(d2) < xi lies in the 27 surrounding cells.
    // The 27 first-probe loads are issued together (independent addresses) before any of them is consumed.
    {
        unsigned long long want[27], got[27];
        unsigned int slot[27];
#pragma unroll
        for (int q = 0; q < 27; ++q) {
            want[q] = pack_ikey(gx + (q / 9 - 1), gy + ((q / 3) % 3 - 1), gz + (q % 3 - 1));
            slot[q] = hash_key(want[q]) & M.gmask;
        }
#pragma unroll
        for (int q = 0; q < 27; ++q) got[q] = M.gkeys[slot[q]];
        int vid[27];
#pragma unroll
        for (int q = 0; q < 27; ++q) {
            int s = -1;
            if (got[q] == want[q]) s = (int)slot[q];
            else if (got[q] != IM_EMPTY_KEY) s = table_find(M.gkeys, M.gmask, want[q]);  // collision chain: rare
            vid[q] = s >= 0 ? M.gval[s] : -1;
        }
        IM_STAMP(58, vid[0] + vid[13] + vid[26]);
        bool close = false;
#pragma unroll
        for (int q = 0; q < 27; ++q) {
            if (vid[q] < 0) continue;
            const float4 v = M.vpos[vid[q]];
            if ((double)sqrtf(dist2f(px, py, pz, v.x, v.y, v.z)) < P.xi) close = true;
        }
        IM_STAMP(59, close);
        if (close) return;
    }
    // survives the old map: enters the per-frame candidate grid, decided in cand_resolve
    F.cand_status[c] = CAND_UNDECIDED;
    int cc = 0;
    const int cs = table_insert(F.ckeys, F.cmask, gkey, &cc);
    if (cs < 0) { im_atomic_or(&M.cnt[3], IM_MERR_HASH_FULL); F.cand_status[c] = CAND_REJECT; return; }
    // push onto the cell's candidate list
#if defined(__CUDA_ARCH__)
    // The lists are only walked by the kernels after this one, so the moment in which the head already names c while c's link is not
    // stored yet is never observed: one exchange, no fence, no retry.  (A next-store / fence / compare-and-swap loop stood here; the
    // stamps put 28-30 k of this function's ~45 k cycles into it, profiles/mstamps_r02y.txt.)
    F.cand_next[c] = atomicExch(&F.chead[cs], c);
#else
    F.cand_next[c] = F.chead[cs];
    F.chead[cs] = c;
#endif
    IM_STAMP(60, 0);
}

// earlier candidates that conflict with c: same xi-cell (:473-481 on a vertex accepted earlier in this frame) or
// closer than xi (:507-517).  Fills cand_conf (or marks overflow).
IM_HDN inline void cand_conflicts(const MeshDev& M, const MeshParams& P, const FrameBuf& F, int c) {
    if (F.cand_status[c] != CAND_UNDECIDED) return;
    const float* p = F.pts + (size_t)c * F.step * 3;
    int gx, gy, gz;
    unpack_ikey(F.cand_gkey[c], &gx, &gy, &gz);
    int n = 0;
    bool overflow = false;
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dz = -1; dz <= 1; ++dz) {
                const int cs = table_find(F.ckeys, F.cmask, pack_ikey(gx + dx, gy + dy, gz + dz));
                if (cs < 0) continue;
                const bool same = (dx == 0 && dy == 0 && dz == 0);
                for (int j = F.chead[cs]; j >= 0; j = F.cand_next[j]) {
                    if (j >= c) continue;
                    const float* q = F.pts + (size_t)j * F.step * 3;
                    if (same || (double)sqrtf(dist2f(p[0], p[1], p[2], q[0], q[1], q[2])) < P.xi) {
                        if (n < IM_CONF_K) F.cand_conf[(size_t)c * IM_CONF_K + n] = j;
                        else overflow = true;
                        ++n;
                    }
                }
            }
    F.cand_nconf[c] = overflow ? -1 : n;
}

// one polling pass for candidate c: accepted iff every earlier conflicting candidate is rejected, rejected iff one
// of them is accepted.  Returns true when decided.  (Sequential greedy semantics of :464-545.)
IM_HDN inline bool cand_poll(const MeshDev& M, const MeshParams& P, const FrameBuf& F, int c) {
    bool pending = false;
    const int nc = F.cand_nconf[c];
    if (nc >= 0) {
        for (int k = 0; k < nc; ++k) {
            const int s = im_vload(&F.cand_status[F.cand_conf[(size_t)c * IM_CONF_K + k]]);
            if (s == CAND_ACCEPT) { F.cand_status[c] = CAND_REJECT; return true; }
            if (s == CAND_UNDECIDED) pending = true;
        }
    } else {
        const float* p = F.pts + (size_t)c * F.step * 3;
        int gx, gy, gz;
        unpack_ikey(F.cand_gkey[c], &gx, &gy, &gz);
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dz = -1; dz <= 1; ++dz) {
                    const int cs = table_find(F.ckeys, F.cmask, pack_ikey(gx + dx, gy + dy, gz + dz));
                    if (cs < 0) continue;
                    const bool same = (dx == 0 && dy == 0 && dz == 0);
                    for (int j = F.chead[cs]; j >= 0; j = F.cand_next[j]) {
                        if (j >= c) continue;
                        const float* q = F.pts + (size_t)j * F.step * 3;
                        if (same || (double)sqrtf(dist2f(p[0], p[1], p[2], q[0], q[1], q[2])) < P.xi) {
                            const int s = im_vload(&F.cand_status[j]);
                            if (s == CAND_ACCEPT) { F.cand_status[c] = CAND_REJECT; return true; }
                            if (s == CAND_UNDECIDED) pending = true;
                        }
                    }
                }
    }
    if (pending) return false;
    F.cand_status[c] = CAND_ACCEPT;
    return true;
}

// accepted candidate -> vertex (pointcloud_rgbd.cpp:518-540); id = vertices before the frame + accepted before c
IM_HDN inline void cand_commit(const MeshDev& M, const MeshParams& P, const FrameBuf& F, int c, int base) {
    if (F.cand_status[c] != CAND_ACCEPT) return;
    const int id = base + F.cand_scan[c];
    if (id >= M.max_v) { im_atomic_or(&M.cnt[3], IM_MERR_VERT_POOL); return; }
    const float* p = F.pts + (size_t)c * F.step * 3;
    M.vpos[id] = make_float4(p[0], p[1], p[2], 0.f);
    M.vsmooth[(size_t)id * 3 + 0] = (double)p[0];
    M.vsmooth[(size_t)id * 3 + 1] = (double)p[1];
    M.vsmooth[(size_t)id * 3 + 2] = (double)p[2];
    M.v_tri_head[id] = -1;
    int created = 0;
    const int gs = table_insert(M.gkeys, M.gmask, F.cand_gkey[c], &created);
    if (gs < 0) { im_atomic_or(&M.cnt[3], IM_MERR_HASH_FULL); return; }
    M.gval[gs] = id;
    const int vs = F.cand_vslot[c];
    const int pos = im_atomic_add(&M.vox_count[vs], 1);
    F.cand_pos[c] = pos;
    if (pos >= IM_VCHUNKS * 16) { im_atomic_or(&M.cnt[3], IM_MERR_VOXEL_CAP); F.cand_pos[c] = -1; }
    else if ((pos & 15) == 0) {   // first vertex of a chunk: allocate it (published for the placement kernel that follows)
        const int ch = im_atomic_add(&M.cnt[27], 1);
        if (ch >= M.max_vchunks) { im_atomic_or(&M.cnt[3], IM_MERR_VERT_POOL); F.cand_pos[c] = -1; }
        else M.vox_chunk[(size_t)vs * IM_VCHUNKS + (pos >> 4)] = ch;
    }
    im_atomic_add(&M.vox_new_added[vs], 1);
    M.vox_meshing_times[vs] = 0;
    im_atomic_add(&M.cnt[10], 1);
}

// second half of the append: the vertex goes into its voxel's chunk (the chunk table is complete after cand_commit)
IM_HDN inline void cand_place(const MeshDev& M, const FrameBuf& F, int c, int base) {
    if (F.cand_status[c] != CAND_ACCEPT || F.cand_pos[c] < 0) return;
    const int id = base + F.cand_scan[c];
    if (id >= M.max_v) return;
    const int vs = F.cand_vslot[c], pos = F.cand_pos[c];
    const int ch = M.vox_chunk[(size_t)vs * IM_VCHUNKS + (pos >> 4)];
    const float* p = F.pts + (size_t)c * F.step * 3;
    float4 v;
    v.x = p[0]; v.y = p[1]; v.z = p[2];
    int idb = id;
    memcpy(&v.w, &idb, 4);
    M.vchunk_pts[(size_t)ch * 16 + (pos & 15)] = v;
}

// activation test of ImMesh_mesh_reconstruction.cpp:132-151 (thread per activated voxel)
IM_HDN inline void voxel_select(const MeshDev& M, const FrameBuf& F, int a) {
    const int vs = F.act[a];
    if (M.vox_meshing_times[vs] >= 1 || M.vox_new_added[vs] < 0) return;
    M.vox_meshing_times[vs] += 1;
    M.vox_new_added[vs] = 0;
    if (M.vox_count[vs] < 3) return;
    if (F.shard_n > 1 && voxel_owner(M.vkeys[vs], F.shard_n) != F.shard_rank) return;   // another rank meshes this voxel
    // two-ended work list: populous voxels (the expensive ones) from the front, the rest from the back, so that the
    // per-voxel stages start the long jobs first (cnt[6] = front count, cnt[17] = back count)
    int w;
    if (M.vox_count[vs] >= 20) w = im_atomic_add(&M.cnt[6], 1);
    else w = F.max_work - 1 - im_atomic_add(&M.cnt[17], 1);
    if (w >= 0 && w < F.max_work && M.cnt[6] + M.cnt[17] <= F.max_work) F.work[w] = vs; else { im_atomic_or(&M.cnt[3], IM_MERR_LIST_CAP); return; }
    // dilation items: groups of up to 8 of the voxel's vertices (the kNN queries), so that a populous voxel is spread over
    // several thread blocks instead of serialising on one
    int nq = M.vox_count[vs];
    if (nq > IM_MAXIN) nq = IM_MAXIN;
    const int ng = (nq + 7) >> 3;
    const int base = im_atomic_add(&M.cnt[29], ng);
    if (base + ng > F.max_ditem) { im_atomic_or(&M.cnt[3], IM_MERR_LIST_CAP); return; }
    for (int g = 0; g < ng; ++g) F.ditem[base + g] = (w << 5) | g;
}

// i-th work item (0 <= i < cnt[6] + cnt[17]) -> slot in the two-ended work arrays
IM_HD int work_slot(const MeshDev& M, const FrameBuf& F, int i) {
    const int na = M.cnt[6];
    return i < na ? i : F.max_work - 1 - (i - na);
}
IM_HD int work_total(const MeshDev& M, const FrameBuf& F) {
    const int t = M.cnt[6] + M.cnt[17];
    return t < F.max_work ? t : F.max_work;
}

// ------------------------------------------------------------------ exact integer predicates
IM_HD int orient2d_i(int ax, int ay, int bx, int by, int cx, int cy) {
    const long long d = (long long)(bx - ax) * (long long)(cy - ay) - (long long)(by - ay) * (long long)(cx - ax);
    return (d > 0) - (d < 0);
}
// > 0 iff d strictly inside the circumcircle of the counter-clockwise triangle (a,b,c).
// double filter on exactly representable integers, exact __int128 fallback.
IM_HDN inline int incircle_i(int ax, int ay, int bx, int by, int cx, int cy, int dx, int dy) {
    const long long adx = (long long)ax - dx, ady = (long long)ay - dy, bdx = (long long)bx - dx, bdy = (long long)by - dy, cdx = (long long)cx - dx, cdy = (long long)cy - dy;
    const long long al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
    const long long m1 = bdx * cdy - bdy * cdx, m2 = cdx * ady - cdy * adx, m3 = adx * bdy - ady * bdx;
    // |coordinates| < 2^25 => lifts and minors < 2^52: exact in double; three products are rounded once each
    const double t1 = (double)al * (double)m1, t2 = (double)bl * (double)m2, t3 = (double)cl * (double)m3;
    const double det = (t1 + t2) + t3;
    const double bound = ((fabs(t1) + fabs(t2)) + fabs(t3)) * 1.0e-15;
    if (det > bound) return 1;
    if (det < -bound) return -1;
    const __int128 ex = (__int128)al * (__int128)m1 + (__int128)bl * (__int128)m2 + (__int128)cl * (__int128)m3;
    return (ex > 0) - (ex < 0);
}

struct DTri { short v[3]; short alive; };  // local vertex indices, -1 = ghost vertex
#define IM_GHOST (-1)

// conflict test of the Bowyer-Watson step: finite triangle -> strict in-circle; ghost (u,v,inf) -> strictly left of
// u->v, or on the open segment uv
IM_HDN inline bool dt_conflict(const DTri& t, const int2* P, int p) {
    const int px = P[p].x, py = P[p].y;
    for (int i = 0; i < 3; ++i)
        if (t.v[i] == IM_GHOST) {
            const int u = t.v[(i + 1) % 3], v = t.v[(i + 2) % 3];
            const int o = orient2d_i(P[u].x, P[u].y, P[v].x, P[v].y, px, py);
            if (o > 0) return true;
            if (o < 0) return false;
            const long long dx = (long long)P[v].x - P[u].x, dy = (long long)P[v].y - P[u].y;
            const long long t1 = ((long long)px - P[u].x) * dx + ((long long)py - P[u].y) * dy;
            const long long t2 = ((long long)px - P[v].x) * dx + ((long long)py - P[v].y) * dy;
            return t1 > 0 && t2 < 0;
        }
    return incircle_i(P[t.v[0]].x, P[t.v[0]].y, P[t.v[1]].x, P[t.v[1]].y, P[t.v[2]].x, P[t.v[2]].y, px, py) > 0;
}

// float circumcircle (centre, padded squared radius) used only to prune the exact in-circle tests
IM_HD void circumcircle_f(const int2* P, int a, int b, int c, float* out) {
    const double bx = (double)(P[b].x - P[a].x), by = (double)(P[b].y - P[a].y);
    const double cx = (double)(P[c].x - P[a].x), cy = (double)(P[c].y - P[a].y);
    const double d = 2.0 * (bx * cy - by * cx);
    const double b2 = bx * bx + by * by, c2 = cx * cx + cy * cy;
    // one float reciprocal refined by a Newton step instead of two double divisions (this is only a pruning filter;
    // the 0.2 % padding below dwarfs the ~1e-14 relative error); d == 0 gives inf/nan -> the guard below disables the filter
    double inv = (double)(1.0f / (float)d);
    inv = inv * (2.0 - d * inv);
    const double ux = (cy * b2 - by * c2) * inv, uy = (bx * c2 - cx * b2) * inv;
    const double r2 = ux * ux + uy * uy;
    const double ccx = (double)P[a].x + ux, ccy = (double)P[a].y + uy;
    if (!(fabs(ccx) < 1.0e9) || !(fabs(ccy) < 1.0e9) || !(r2 < 1.0e18)) { out[0] = 0.f; out[1] = 0.f; out[2] = INFINITY; return; }
    out[0] = (float)ccx; out[1] = (float)ccy;
    out[2] = (float)(r2 * 1.002) + 64.0f;   // conservative: float rounding of centre/radius is ~1e-6 relative
}

// Block-cooperative Delaunay triangulation of n snapped points (Bowyer-Watson, exact predicates, insertion in
// index order after a non-degenerate seed).  tris/ntri: triangle pool in shared memory.  scratch: >= 784 ints
// (scratch[4] = seed found, scratch[6] = capacity overflow).
// Returns (via *ntri) the pool size; dead slots have alive == 0.  ok=false when all points are collinear.
IM_HDN inline void delaunay_block(const int2* P, int n, DTri* tris, int max_tris, int* ntri, int* scratch, int tid, int nthreads, float (*circ)[3] = nullptr) {
    int* s_cav_n = scratch;        // [0]
    int* s_edge_n = scratch + 1;   // [1]
    int* s_seed = scratch + 2;     // [2..4] i1, i2, ok
    int* s_new_n = scratch + 5;
    int* s_ovf = scratch + 6;      // [6] cavity / pool overflow flag
    int* s_cav = scratch + 8;      // [256] cavity triangle indices
    int* s_edges = scratch + 264;  // [260*2] boundary edges (a,b)
    if (tid == 0) {
        *s_ovf = 0;
        int i1 = -1, i2 = -1;
        for (int i = 1; i < n; ++i)
            if (P[i].x != P[0].x || P[i].y != P[0].y) { i1 = i; break; }
        if (i1 >= 0)
            for (int i = 1; i < n; ++i)
                if (i != i1 && orient2d_i(P[0].x, P[0].y, P[i1].x, P[i1].y, P[i].x, P[i].y) != 0) { i2 = i; break; }
        s_seed[0] = i1; s_seed[1] = i2; s_seed[2] = (i1 >= 0 && i2 >= 0) ? 1 : 0;
        *ntri = 0;
        if (s_seed[2]) {
            int a = 0, b = i1, c = i2;
            if (orient2d_i(P[a].x, P[a].y, P[b].x, P[b].y, P[c].x, P[c].y) < 0) { const int t = b; b = c; c = t; }
            const short tv[4][3] = {{(short)a, (short)b, (short)c}, {(short)c, (short)b, IM_GHOST}, {(short)a, (short)c, IM_GHOST}, {(short)b, (short)a, IM_GHOST}};
            for (int k = 0; k < 4; ++k) { tris[k].v[0] = tv[k][0]; tris[k].v[1] = tv[k][1]; tris[k].v[2] = tv[k][2]; tris[k].alive = 1; }
            if (circ) circumcircle_f(P, a, b, c, circ[0]);
            *ntri = 4;
        }
    }
    IM_SYNCBLOCK_M();
    if (!s_seed[2]) return;
    const int i1 = s_seed[0], i2 = s_seed[1];
    for (int p = 1; p < n; ++p) {
        if (p == i1 || p == i2) continue;
        if (tid == 0) { *s_cav_n = 0; *s_edge_n = 0; *s_new_n = 0; }
        IM_SYNCBLOCK_M();
        const int nt = *ntri;
        const float pxf = (float)P[p].x, pyf = (float)P[p].y;
        for (int t = tid; t < nt; t += nthreads) {
            if (!tris[t].alive) continue;
            if (circ && tris[t].v[0] != IM_GHOST && tris[t].v[1] != IM_GHOST && tris[t].v[2] != IM_GHOST) {
                const float dx = pxf - circ[t][0], dy = pyf - circ[t][1];
                if (dx * dx + dy * dy > circ[t][2]) continue;   // certainly outside the circumcircle
            }
            if (dt_conflict(tris[t], P, p)) {
                const int k = im_atomic_add(s_cav_n, 1);
                if (k < 256) s_cav[k] = t;
            }
        }
        IM_SYNCBLOCK_M();
        int nc = *s_cav_n;
        if (nc > 256) { nc = 256; *s_ovf = 1; }
        // boundary edges: directed edge (a,b) of a cavity triangle whose reverse is not a cavity edge
        for (int e = tid; e < nc * 3; e += nthreads) {
            const DTri& t = tris[s_cav[e / 3]];
            const int a = t.v[(e % 3 + 1) % 3], b = t.v[(e % 3 + 2) % 3];
            bool interior = false;
            for (int f = 0; f < nc * 3 && !interior; ++f) {
                const DTri& u = tris[s_cav[f / 3]];
                if (u.v[(f % 3 + 1) % 3] == b && u.v[(f % 3 + 2) % 3] == a) interior = true;
            }
            if (!interior) {
                const int k = im_atomic_add(s_edge_n, 1);
                if (k < 260) { s_edges[2 * k] = a; s_edges[2 * k + 1] = b; }
            }
        }
        IM_SYNCBLOCK_M();
        int ne = *s_edge_n;
        if (ne > 260) { ne = 260; *s_ovf = 1; }
        // kill the cavity, write (a,b,p) per boundary edge: reuse cavity slots first, then grow the pool
        for (int k = tid; k < nc; k += nthreads) tris[s_cav[k]].alive = 0;
        IM_SYNCBLOCK_M();
        for (int k = tid; k < ne; k += nthreads) {
            const int slot = (k < nc) ? s_cav[k] : (nt + (k - nc));
            if (slot < max_tris) {
                tris[slot].v[0] = (short)s_edges[2 * k];
                tris[slot].v[1] = (short)s_edges[2 * k + 1];
                tris[slot].v[2] = (short)p;
                tris[slot].alive = 1;
                if (circ && s_edges[2 * k] != IM_GHOST && s_edges[2 * k + 1] != IM_GHOST) circumcircle_f(P, s_edges[2 * k], s_edges[2 * k + 1], p, circ[slot]);
            }
        }
        if (tid == 0 && ne > nc) {
            if (nt + (ne - nc) <= max_tris) *ntri = nt + (ne - nc);
            else { *ntri = max_tris; *s_ovf = 1; }
        }
        IM_SYNCBLOCK_M();
    }
}

// smallest double c with !(acos(c) * 57.3 > 150) under glibc: a face is dropped iff cos < this (and cos >= -1)
#define IM_COS150 (-8.65928972248464878803e-01)
IM_HD bool angle_bad(double ax, double ay, double bx, double by, double cx, double cy) {  // compute_angle at apex a
    const double abx = bx - ax, aby = by - ay, acx = cx - ax, acy = cy - ay;
    const double cosv = (abx * acx + aby * acy) / (sqrt(abx * abx + aby * aby) * sqrt(acx * acx + acy * acy));
    return cosv < IM_COS150 && cosv >= -1.0;
}

// correct_triangle_index, mesh_rec_geometry.cpp:399-433 (smoothed vertex positions)
IM_HDN inline int compute_flip(const MeshDev& M, int ia, int ib, int ic, const double* cam, const double* short_axis) {
    const double* a = M.vsmooth + (size_t)ia * 3;
    const double* b = M.vsmooth + (size_t)ib * 3;
    const double* c = M.vsmooth + (size_t)ic * 3;
    const double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    const double ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const double tc[3] = {cam[0] - a[0], cam[1] - a[1], cam[2] - a[2]};
    double nrm[3] = {ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]};
    const double nn = sqrt(dot3(nrm, nrm));
    if (nn != 0) { nrm[0] = nrm[0] / nn; nrm[1] = nrm[1] / nn; nrm[2] = nrm[2] / nn; }
    else { nrm[0] = 0; nrm[1] = 0; nrm[2] = 1; }
    double sa[3] = {short_axis[0], short_axis[1], short_axis[2]};
    if (dot3(sa, tc) < 0) { sa[0] = -sa[0]; sa[1] = -sa[1]; sa[2] = -sa[2]; }
    return (dot3(sa, nrm) < 0) ? 0 : 1;
}

// ------------------------------------------------------------------ triangle store
IM_HD unsigned int tri_hash(int a, int b, int c) {
    unsigned long long k = ((unsigned long long)(unsigned int)a * 0x9E3779B97F4A7C15ull) ^ ((unsigned long long)(unsigned int)b * 0xC2B2AE3D27D4EB4Full) ^
                           ((unsigned long long)(unsigned int)c * 0x165667B19E3779F9ull);
    k ^= k >> 31; k *= 0xff51afd7ed558ccdull; k ^= k >> 33;
    return (unsigned int)k;
}
IM_HDN inline int tri_find(const MeshDev& M, int a, int b, int c) {
    unsigned int s = tri_hash(a, b, c) & M.tmask;
    for (unsigned int probe = 0; probe <= M.tmask; ++probe) {
        const int t = im_vload(&M.thash[s]);
        if (t < 0) return -1;
        const int4 r = M.tri[t];
        if (r.x == a && r.y == b && r.z == c) return t;
        s = (s + 1) & M.tmask;
    }
    return -1;
}
// insert_triangle (triangle.hpp:330-395): find-or-create, mark alive, link into the three incidence lists on creation
IM_HDN inline void tri_add(const MeshDev& M, int a, int b, int c, unsigned long long flipword) {
    unsigned int s = tri_hash(a, b, c) & M.tmask;
    int mine = -1;
    int t = -1;
    for (unsigned int probe = 0; probe <= M.tmask; ++probe) {
        int cur = im_vload(&M.thash[s]);
        if (cur < 0) {
            if (mine < 0) {
                mine = im_atomic_add(&M.cnt[1], 1);
                if (mine >= M.max_t) { im_atomic_or(&M.cnt[3], IM_MERR_TRI_POOL); return; }
                M.tri[mine] = make_int4(a, b, c, 0);
                M.tri_flip[mine] = 0ull;
                M.tri_next[(size_t)mine * 3 + 0] = -1; M.tri_next[(size_t)mine * 3 + 1] = -1; M.tri_next[(size_t)mine * 3 + 2] = -1;
                im_fence();
            }
            cur = im_atomic_cas32(&M.thash[s], -1, mine);
            if (cur < 0) { t = mine; break; }
        }
        const int4 r = M.tri[cur];
        if (r.x == a && r.y == b && r.z == c) { t = cur; break; }
        s = (s + 1) & M.tmask;
    }
    if (t < 0) { im_atomic_or(&M.cnt[3], IM_MERR_HASH_FULL); return; }
    if (t == mine) {
        const int vs[3] = {a, b, c};
#if defined(__CUDA_ARCH__)
        // three independent list insertions.  The incidence lists are walked by other kernels only (pull, remove: before or after
        // this one in stream order), so head-before-link is never observed: one exchange per list, no fence, no retry loop.
        int prev[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) prev[k] = atomicExch(&M.v_tri_head[vs[k]], t);
#pragma unroll
        for (int k = 0; k < 3; ++k) M.tri_next[(size_t)t * 3 + k] = prev[k];
#else
        for (int k = 0; k < 3; ++k) {
            M.tri_next[(size_t)t * 3 + k] = M.v_tri_head[vs[k]];
            M.v_tri_head[vs[k]] = t;
        }
#endif
    }
    // (a record allocated by a thread that lost the publication race stays unused: a leaked slot, never linked)
    if (im_atomic_exch(&M.tri[t].w, 1) == 0) im_atomic_add(&M.cnt[2], 1);
    im_atomic_max64(&M.tri_flip[t], flipword);
}
IM_HD void tri_remove(const MeshDev& M, int t) {  // erase_triangle, triangle.hpp:164-210
    if (im_atomic_exch(&M.tri[t].w, 0) == 1) im_atomic_add(&M.cnt[2], -1);
}

}  // namespace immesh
