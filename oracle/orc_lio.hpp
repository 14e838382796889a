// ORACLE (test infrastructure, NOT product code) -- CPU restatement of ImMesh's per-scan
// localization hot path: VoxelMap (hash of adaptive octrees with probabilistic planes),
// point-to-plane residual selection, IESKF update and map growth.
// Parity status: "parity unpinned" (see orc_math.hpp header).  All file:line citations are
// relative to /root/reference.
#pragma once
#include <algorithm>
#include <cstdio>
#include <unordered_map>
#include <vector>

#include "orc_math.hpp"

namespace orc {

struct LioCfg {
    double voxel_size;          // voxel/voxel_size  -> m_max_voxel_size (voxel_mapping.hpp:176)
    int max_layer;              // voxel/max_layer
    int layer_init_size[5];     // voxel/layer_init_size
    int max_points_size;        // voxel/max_points_size
    double min_eigen_value;     // voxel/min_eigen_value (narrowed to float planer_threshold)
    double dept_err, beam_err;  // narrowed to float by calcBodyVar's signature
    double extR[9], extT[3];    // LiDAR->IMU extrinsic (m_extR, m_extT)
    int max_iteration;          // NUM_MAX_ITERATIONS
    int calib_laser;            // preprocess/calib_laser (weights use CALIB_ANGLE_COV, :1513)
    int sum_mode;               // 0: fixed-point order-free sums (matches the CUDA product bit for bit)
                                // 1: plain serial double sums (closest to the reference's Eigen GEMM)
    int omp_threads;            // threads for the residual loop (reference: MP_PROC_NUM = 4)
    int plane_var_mode;         // 0: plane covariance from running moments (matches the CUDA product); 1: the reference's per-point loop (voxel_loc.cpp:76-121)
    int solve_mode;             // 0: 6x6 matrix-inversion-lemma form of the IESKF gain (matches the CUDA product); 1: the reference's two 18x18 inverses
};

struct PV {  // Point_with_var, voxel_loc.hpp:75-80
    double pb[3];   // m_point (body frame on the residual path; WORLD frame on the insert path)
    double pw[3];   // m_point_world
    double var[6];  // m_var, symmetric
};

struct Plane {  // voxel_loc.hpp:89-104 (fields that reach an output)
    double center[3] = {0, 0, 0};
    double normal[3] = {0, 0, 0};
    double plane_var[21];  // upper triangle of the 6x6, row-major: (0,0),(0,1)..(0,5),(1,1)..
    float radius = 0;
    float min_eigen_value = 1;
    float d = 0;
    int points_size = 0;
    bool is_plane = false;
    bool is_init = false;
    int id = 0;
    Plane() { for (double& v : plane_var) v = 0; }
};

inline int pv21_idx(int i, int j) {  // i <= j
    return i * 6 - (i * (i - 1)) / 2 + (j - i);
}

struct OctoTree {  // voxel_loc.hpp:129-177
    std::vector<PV> temp_points;
    Plane plane;
    int layer;
    int octo_state = 0;
    OctoTree* leaves[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double voxel_center[3] = {0, 0, 0};
    float quater_length = 0;
    int new_points = 0;
    bool init_octo = false;
    bool update_enable = true;
    explicit OctoTree(int layer_) : layer(layer_) {}
    ~OctoTree() { for (auto* l : leaves) delete l; }
};

struct VoxelKey {
    int64_t x, y, z;
    bool operator==(const VoxelKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct VoxelKeyHash {  // voxel_loc.hpp:118-126 (HASH_P 116101, MAX_N 1e10)
    size_t operator()(const VoxelKey& s) const {
        return (size_t)((((s.z) * 116101LL) % 10000000000LL + (s.y)) * 116101LL % 10000000000LL + (s.x));
    }
};

struct Ptpl {  // voxel_loc.hpp:63-73
    double point[3], normal[3], center[3];
    double plane_var[21];
    int layer;
    double d;
    int src_index;  // index into the down-sampled scan (not in the reference struct; for parity tests)
};

struct State {  // StatesGroup, common_lib.h:199-288
    double rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double pos[3] = {0, 0, 0}, vel[3] = {0, 0, 0}, bg[3] = {0, 0, 0}, ba[3] = {0, 0, 0}, grav[3] = {0, 0, 0};
    double cov[324];
    State() {
        for (int i = 0; i < 324; ++i) cov[i] = 0;
        for (int i = 0; i < 18; ++i) cov[i * 18 + i] = 0.0000001;  // INIT_COV
    }
};

struct IterStats {
    double HTH[36];  // full symmetric 6x6
    double HTz[6];
    int n_match;
    double total_residual;
    double solution[18];
    int converged;
};

// key rule, voxel_mapping.cpp:118-127 (insert: float voxel_size) and :172-181 (lookup: double)
inline VoxelKey voxel_key(const double* p, double vs) {
    int64_t k[3];
    for (int j = 0; j < 3; ++j) {
        float loc = (float)(p[j] / vs);
        if (loc < 0) loc = (float)((double)loc - 1.0);
        k[j] = (int64_t)loc;
    }
    return VoxelKey{k[0], k[1], k[2]};
}

class LioOracle {
  public:
    LioCfg cfg;
    State state;
    std::unordered_map<VoxelKey, OctoTree*, VoxelKeyHash> feat_map;
    int g_plane_id = 0;
    float planer_threshold;
    float dept_err_f, beam_err_f;
    double dir_var;        // pow(sin(DEG2RAD(beam_err)),2)
    double dir_var_calib;  // same with CALIB_ANGLE_COV 0.01 (common_lib.h:41)
    // per-scan caches (m_body_cov_list / m_cross_mat_list, voxel_mapping.cpp:1302-1316)
    std::vector<double> body_cov;   // N x 6
    std::vector<double> cross_mat;  // N x 9
    std::vector<IterStats> iter_stats;
    std::vector<int> last_match_idx, last_match_layer;  // of the final residual build
    int iters_run = 0;

    explicit LioOracle(const LioCfg& c) : cfg(c) {
        planer_threshold = (float)cfg.min_eigen_value;
        dept_err_f = (float)cfg.dept_err;
        beam_err_f = (float)cfg.beam_err;
        // DEG2RAD is PCL's macro ((x)*0.017453293), pcl/common/angles.h (PCL is not vendored)
        const double s = std::sin((double)beam_err_f * 0.017453293);
        dir_var = s * s;
        const double sc = std::sin((double)(float)0.01 * 0.017453293);
        dir_var_calib = sc * sc;
    }
    ~LioOracle() { for (auto& kv : feat_map) delete kv.second; }

    // ------------------------------------------------------------------ calcBodyVar, voxel_mapping.cpp:1221-1241
    static void calc_body_var(double* pb, float range_inc, double direction_var, double* var6) {
        if (pb[2] == 0) pb[2] = 0.0001;
        const float range = (float)std::sqrt((pb[0] * pb[0] + pb[1] * pb[1]) + pb[2] * pb[2]);
        const float range_var = range_inc * range_inc;
        const double nrm = std::sqrt((pb[0] * pb[0] + pb[1] * pb[1]) + pb[2] * pb[2]);
        const double dir[3] = {pb[0] / nrm, pb[1] / nrm, pb[2] / nrm};
        double hat[9];
        skew(dir, hat);
        double b1[3] = {1.0, 1.0, -(dir[0] + dir[1]) / dir[2]};
        const double n1 = std::sqrt((b1[0] * b1[0] + b1[1] * b1[1]) + b1[2] * b1[2]);
        b1[0] = b1[0] / n1; b1[1] = b1[1] / n1; b1[2] = b1[2] / n1;
        double b2[3] = {b1[1] * dir[2] - b1[2] * dir[1], b1[2] * dir[0] - b1[0] * dir[2], b1[0] * dir[1] - b1[1] * dir[0]};
        const double n2 = std::sqrt((b2[0] * b2[0] + b2[1] * b2[1]) + b2[2] * b2[2]);
        b2[0] = b2[0] / n2; b2[1] = b2[1] / n2; b2[2] = b2[2] / n2;
        // A = range * direction_hat * N,  N = [b1 b2]
        double A[3][2];
        const double rg = (double)range;
        for (int i = 0; i < 3; ++i) {
            const double h0 = rg * hat[i * 3 + 0], h1 = rg * hat[i * 3 + 1], h2 = rg * hat[i * 3 + 2];
            A[i][0] = (h0 * b1[0] + h1 * b1[1]) + h2 * b1[2];
            A[i][1] = (h0 * b2[0] + h1 * b2[1]) + h2 * b2[2];
        }
        const double rv = (double)range_var;
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j)
                var6[sym6_idx(i, j)] = (dir[i] * rv) * dir[j] + ((A[i][0] * direction_var) * A[j][0] + (A[i][1] * direction_var) * A[j][1]);
    }

    // p_w = (float)( R (R_ext p + t_ext) + t ),  transformLidar, voxel_mapping_common.cpp:709-726
    void body_to_world_f(const double* R, const double* t, const double* pb, double* pw_as_double) const {
        double q[3], w[3];
        mat3_vec(cfg.extR, pb, q);
        for (int i = 0; i < 3; ++i) q[i] = q[i] + cfg.extT[i];
        mat3_vec(R, q, w);
        for (int i = 0; i < 3; ++i) pw_as_double[i] = (double)(float)(w[i] + t[i]);
    }
    void body_to_world_d(const double* R, const double* t, const double* pb, double* pw) const {  // pointBodyToWorld, :121-131
        double q[3], w[3];
        mat3_vec(cfg.extR, pb, q);
        for (int i = 0; i < 3; ++i) q[i] = q[i] + cfg.extT[i];
        mat3_vec(R, q, w);
        for (int i = 0; i < 3; ++i) pw[i] = w[i] + t[i];
    }

    // ------------------------------------------------------------------ init_plane, voxel_loc.cpp:47-139
    void init_plane(const std::vector<PV>& pts, Plane* pl, const double* vc) {
        for (double& v : pl->plane_var) v = 0;
        double cov[6] = {0, 0, 0, 0, 0, 0}, c[3] = {0, 0, 0};
        pl->normal[0] = pl->normal[1] = pl->normal[2] = 0;
        const int n = (int)pts.size();
        pl->points_size = n;
        pl->radius = 0;
        for (const PV& pv : pts) {
            const double* p = pv.pb;
            cov[0] += p[0] * p[0]; cov[1] += p[0] * p[1]; cov[2] += p[0] * p[2];
            cov[3] += p[1] * p[1]; cov[4] += p[1] * p[2]; cov[5] += p[2] * p[2];
            c[0] += p[0]; c[1] += p[1]; c[2] += p[2];
        }
        const double dn = (double)n;
        for (int i = 0; i < 3; ++i) c[i] = c[i] / dn;
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j) cov[sym6_idx(i, j)] = cov[sym6_idx(i, j)] / dn - c[i] * c[j];
        for (int i = 0; i < 3; ++i) pl->center[i] = c[i];
        double ev[3], U[9];
        jacobi_eig3(cov, ev, U);
        int imin = 0, imax = 0;
        for (int i = 1; i < 3; ++i) {
            if (ev[i] < ev[imin]) imin = i;
            if (ev[i] > ev[imax]) imax = i;
        }
        const int imid = 3 - imin - imax;
        (void)imid;
        if (ev[imin] < (double)planer_threshold) {
            const double invn = 1.0 / dn;
            // F_m = (p - c)^T / (n (lambda_min - lambda_m)) * (...)  (voxel_loc.cpp:88-91).  The scalar division is applied as a
            // multiplication by the reciprocal (one division per refit instead of three per point; <= 1 ulp per element).
            double Mm[3][9], sm[3];
            for (int m = 0; m < 3; ++m) {
                if (m == imin) continue;
                sm[m] = 1.0 / (dn * (ev[imin] - ev[m]));
                for (int j = 0; j < 3; ++j)
                    for (int k = 0; k < 3; ++k)
                        Mm[m][j * 3 + k] = U[j * 3 + m] * U[k * 3 + imin] + U[j * 3 + imin] * U[k * 3 + m];
            }
            if (cfg.plane_var_mode == 0) {
                // plane_var_mode 0 (what the CUDA product computes, bit for bit): J_i is linear in d_i = p_i - c, so
                // sum_i J_i Sigma_i J_i^T is a fixed contraction of the moments  Q2_jk = sum q_j q_k Sigma_i,  Q1_j = sum q_j Sigma_i,
                // S0 = sum Sigma_i  about the node's fixed voxel centre (q = p - vc; the product keeps them as running sums in
                // append order -> refits cost O(1)).  With e = c - vc, G_j[a][b] = U[a][m0] (s0 M0[j][b]) + U[a][m1] (s1 M1[j][b]):
                //   W_jk = ((Q2_jk - e_j Q1_k) - e_k Q1_j) + (e_j e_k) S0,  V_j = Q1_j - e_j S0,
                //   top-left = sum_jk G_j W_jk G_k^T,  top-right = (sum_j G_j V_j)/n,  bottom-right = (S0/n)/n.
                // plane_var_mode 1 below is the reference's literal per-point loop; tests bound the gap between the two.
                double mq[60];
                for (double& v : mq) v = 0.0;
                for (int ip = 0; ip < n; ++ip) {
                    const PV& pv = pts[ip];
                    const double q[3] = {pv.pb[0] - vc[0], pv.pb[1] - vc[1], pv.pb[2] - vc[2]};
                    for (int e = 0; e < 60; ++e) {
                        double w;
                        int sidx;
                        if (e < 36) {
                            const int pr = e / 6;
                            sidx = e - pr * 6;
                            const int j = pr < 3 ? 0 : (pr < 5 ? 1 : 2), k = pr < 3 ? pr : (pr < 5 ? pr - 2 : 2);
                            w = q[j] * q[k];
                        } else if (e < 54) {
                            const int j = (e - 36) / 6;
                            sidx = (e - 36) - j * 6;
                            w = q[j];
                        } else {
                            sidx = e - 54;
                            w = 1.0;
                        }
                        mq[e] = mq[e] + w * pv.var[sidx];
                    }
                }
                int m0 = -1, m1 = -1;
                for (int m = 0; m < 3; ++m)
                    if (m != imin) { if (m0 < 0) m0 = m; else m1 = m; }
                const double s0 = sm[m0], s1 = sm[m1];
                const double e3[3] = {c[0] - vc[0], c[1] - vc[1], c[2] - vc[2]};
                for (int ei = 0; ei < 6; ++ei)
                    for (int ej = ei; ej < 6; ++ej) {
                        double r;
                        if (ei >= 3) {
                            r = (invn * mq[54 + sym6_idx(ei - 3, ej - 3)]) * invn;
                        } else {
                            double Ga[9], Gb[9];
                            const int bb = ej < 3 ? ej : 0;
                            for (int j = 0; j < 3; ++j)
                                for (int x = 0; x < 3; ++x) {
                                    Ga[j * 3 + x] = U[ei * 3 + m0] * (s0 * Mm[m0][j * 3 + x]) + U[ei * 3 + m1] * (s1 * Mm[m1][j * 3 + x]);
                                    Gb[j * 3 + x] = U[bb * 3 + m0] * (s0 * Mm[m0][j * 3 + x]) + U[bb * 3 + m1] * (s1 * Mm[m1][j * 3 + x]);
                                }
                            if (ej >= 3) {
                                const int l = ej - 3;
                                double a2 = 0.0;
                                for (int j = 0; j < 3; ++j)
                                    for (int x = 0; x < 3; ++x) {
                                        const int sx = sym6_idx(x, l);
                                        const double v = mq[36 + j * 6 + sx] - e3[j] * mq[54 + sx];
                                        a2 = a2 + Ga[j * 3 + x] * v;
                                    }
                                r = a2 * invn;
                            } else {
                                double a2 = 0.0;
                                for (int j = 0; j < 3; ++j)
                                    for (int k2 = 0; k2 < 3; ++k2) {
                                        const int pr = sym6_idx(j, k2);
                                        const double ejk = e3[j] * e3[k2];
                                        for (int x = 0; x < 3; ++x)
                                            for (int y = 0; y < 3; ++y) {
                                                const int sx = sym6_idx(x, y);
                                                const double w = ((mq[pr * 6 + sx] - e3[j] * mq[36 + k2 * 6 + sx]) - e3[k2] * mq[36 + j * 6 + sx]) + ejk * mq[54 + sx];
                                                a2 = a2 + (Ga[j * 3 + x] * w) * Gb[k2 * 3 + y];
                                            }
                                    }
                                r = a2;
                            }
                        }
                        pl->plane_var[pv21_idx(ei, ej)] = r;
                    }
            }
            for (int ip = 0; ip < n && cfg.plane_var_mode != 0; ++ip) {
                const PV& pv = pts[ip];
                double F[9];
                for (int m = 0; m < 3; ++m) {
                    if (m != imin) {
                        const double v0 = (pv.pb[0] - c[0]) * sm[m], v1 = (pv.pb[1] - c[1]) * sm[m], v2 = (pv.pb[2] - c[2]) * sm[m];
                        for (int k = 0; k < 3; ++k) F[m * 3 + k] = (v0 * Mm[m][0 * 3 + k] + v1 * Mm[m][1 * 3 + k]) + v2 * Mm[m][2 * 3 + k];
                    } else {
                        F[m * 3 + 0] = 0; F[m * 3 + 1] = 0; F[m * 3 + 2] = 0;
                    }
                }
                double A[9], S[9], T[9];
                mat3_mul(U, F, A);  // J top block = evecs * F
                sym6_to_full(pv.var, S);
                mat3_mul(A, S, T);  // T = A * Sigma
                // top-left: A Sigma A^T
                for (int i = 0; i < 3; ++i)
                    for (int j = i; j < 3; ++j)
                        pl->plane_var[pv21_idx(i, j)] += (T[i * 3 + 0] * A[j * 3 + 0] + T[i * 3 + 1] * A[j * 3 + 1]) + T[i * 3 + 2] * A[j * 3 + 2];
                // top-right: (A Sigma) * (I/n)
                for (int i = 0; i < 3; ++i)
                    for (int k = 0; k < 3; ++k) pl->plane_var[pv21_idx(i, 3 + k)] += T[i * 3 + k] * invn;
                // bottom-right: (Sigma/n)/n
                for (int k = 0; k < 3; ++k)
                    for (int l = k; l < 3; ++l) pl->plane_var[pv21_idx(3 + k, 3 + l)] += (invn * S[k * 3 + l]) * invn;
            }
            for (int i = 0; i < 3; ++i) pl->normal[i] = U[i * 3 + imin];
            pl->min_eigen_value = (float)ev[imin];
            pl->radius = (float)std::sqrt(ev[imax]);
            pl->d = (float)(-((pl->normal[0] * c[0] + pl->normal[1] * c[1]) + pl->normal[2] * c[2]));
            pl->is_plane = true;
        } else {
            pl->is_plane = false;
        }
        if (!pl->is_init) {
            pl->id = g_plane_id++;
            pl->is_init = true;
        }
    }

    OctoTree* make_child(OctoTree* parent, const int* xyz) {  // voxel_loc.cpp:186-190 / :279-283
        OctoTree* ch = new OctoTree(parent->layer + 1);
        for (int j = 0; j < 3; ++j)
            ch->voxel_center[j] = parent->voxel_center[j] + (double)((float)(2 * xyz[j] - 1) * parent->quater_length);
        ch->quater_length = parent->quater_length / 2;
        return ch;
    }
    int init_size(const OctoTree* n) const { return cfg.layer_init_size[n->layer]; }

    // voxel_loc.cpp:161-217
    void cut_octo_tree(OctoTree* nd) {
        if (nd->layer >= cfg.max_layer) {
            nd->octo_state = 0;
            return;
        }
        for (size_t i = 0; i < nd->temp_points.size(); ++i) {
            const PV& pv = nd->temp_points[i];
            int xyz[3] = {0, 0, 0};
            for (int j = 0; j < 3; ++j)
                if (pv.pb[j] > nd->voxel_center[j]) xyz[j] = 1;
            const int leaf = 4 * xyz[0] + 2 * xyz[1] + xyz[2];
            if (nd->leaves[leaf] == nullptr) nd->leaves[leaf] = make_child(nd, xyz);
            nd->leaves[leaf]->temp_points.push_back(pv);
            nd->leaves[leaf]->new_points++;
        }
        for (int i = 0; i < 8; ++i) {
            OctoTree* ch = nd->leaves[i];
            if (ch == nullptr) continue;
            if ((int)ch->temp_points.size() > init_size(ch)) {
                init_plane(ch->temp_points, &ch->plane, ch->voxel_center);
                if (ch->plane.is_plane) {
                    ch->octo_state = 0;
                } else {
                    ch->octo_state = 1;
                    cut_octo_tree(ch);
                }
                ch->init_octo = true;
                ch->new_points = 0;
            }
        }
    }
    // voxel_loc.cpp:141-159
    void init_octo_tree(OctoTree* nd) {
        if ((int)nd->temp_points.size() > init_size(nd)) {
            init_plane(nd->temp_points, &nd->plane, nd->voxel_center);
            if (nd->plane.is_plane) {
                nd->octo_state = 0;
            } else {
                nd->octo_state = 1;
                cut_octo_tree(nd);
            }
            nd->init_octo = true;
            nd->new_points = 0;
        }
    }
    // voxel_loc.cpp:219-308
    void update_octo_tree(OctoTree* nd, const PV& pv) {
        if (!nd->init_octo) {
            nd->new_points++;
            nd->temp_points.push_back(pv);
            if ((int)nd->temp_points.size() > init_size(nd)) init_octo_tree(nd);
            return;
        }
        if (nd->plane.is_plane) {
            if (nd->update_enable) {
                nd->new_points++;
                nd->temp_points.push_back(pv);
                if (nd->new_points > 5) {  // m_update_size_threshold_
                    init_plane(nd->temp_points, &nd->plane, nd->voxel_center);
                    nd->new_points = 0;
                }
                if ((int)nd->temp_points.size() >= cfg.max_points_size) {
                    nd->update_enable = false;
                    std::vector<PV>().swap(nd->temp_points);
                    nd->new_points = 0;
                }
            }
            return;
        }
        if (nd->layer < cfg.max_layer) {
            if (!nd->temp_points.empty()) std::vector<PV>().swap(nd->temp_points);
            int xyz[3] = {0, 0, 0};
            for (int j = 0; j < 3; ++j)
                if (pv.pb[j] > nd->voxel_center[j]) xyz[j] = 1;
            const int leaf = 4 * xyz[0] + 2 * xyz[1] + xyz[2];
            if (nd->leaves[leaf] == nullptr) nd->leaves[leaf] = make_child(nd, xyz);
            update_octo_tree(nd->leaves[leaf], pv);
        } else {
            if (nd->update_enable) {
                nd->new_points++;
                nd->temp_points.push_back(pv);
                if (nd->new_points > 5) {
                    init_plane(nd->temp_points, &nd->plane, nd->voxel_center);
                    nd->new_points = 0;
                }
                if ((int)nd->temp_points.size() > 1000) {  // g_max_points, voxel_loc.cpp:45,298
                    nd->update_enable = false;
                    std::vector<PV>().swap(nd->temp_points);
                }
            }
        }
    }

    OctoTree* new_root(const VoxelKey& k) {  // voxel_mapping.cpp:136-141 / :345-350
        const float vsf = (float)cfg.voxel_size;
        OctoTree* t = new OctoTree(0);
        t->quater_length = vsf / 4;
        t->voxel_center[0] = (0.5 + (double)k.x) * (double)vsf;
        t->voxel_center[1] = (0.5 + (double)k.y) * (double)vsf;
        t->voxel_center[2] = (0.5 + (double)k.z) * (double)vsf;
        return t;
    }
    // buildVoxelMap, voxel_mapping.cpp:110-151 (pv.pb carries the WORLD point here)
    void build_voxel_map(const std::vector<PV>& pts) {
        const double vs = (double)(float)cfg.voxel_size;
        for (const PV& pv : pts) {
            const VoxelKey k = voxel_key(pv.pb, vs);
            auto it = feat_map.find(k);
            OctoTree* t;
            if (it != feat_map.end()) {
                t = it->second;
            } else {
                t = new_root(k);
                feat_map[k] = t;
            }
            t->temp_points.push_back(pv);
            t->new_points++;
        }
        for (auto& kv : feat_map) init_octo_tree(kv.second);
    }
    // updateVoxelMap, voxel_mapping.cpp:320-354
    void update_voxel_map(const std::vector<PV>& pts) {
        const double vs = (double)(float)cfg.voxel_size;
        for (const PV& pv : pts) {
            const VoxelKey k = voxel_key(pv.pb, vs);
            auto it = feat_map.find(k);
            OctoTree* t;
            if (it != feat_map.end()) {
                t = it->second;
            } else {
                t = new_root(k);
                feat_map[k] = t;
            }
            update_octo_tree(t, pv);
        }
    }

    // sigma_l = J_nq * plane_var * J_nq^T with J_nq = [p_w - c, -n]  (voxel_mapping.cpp:264-267, :1523-1526)
    static double plane_sigma(const double* pw, const double* center, const double* normal, const double* pv21) {
        const double J[6] = {pw[0] - center[0], pw[1] - center[1], pw[2] - center[2], -normal[0], -normal[1], -normal[2]};
        double acc = 0.0;
        for (int i = 0; i < 6; ++i) {
            double row = 0.0;
            for (int j = 0; j < 6; ++j) {
                const double m = (i <= j) ? pv21[pv21_idx(i, j)] : pv21[pv21_idx(j, i)];
                row = row + m * J[j];
            }
            acc = acc + J[i] * row;
        }
        return acc;
    }
    static double quad_sym6(const double* n, const double* v6) {  // n^T V n
        double acc = 0.0;
        for (int i = 0; i < 3; ++i) {
            const double row = (v6[sym6_idx(i, 0)] * n[0] + v6[sym6_idx(i, 1)] * n[1]) + v6[sym6_idx(i, 2)] * n[2];
            acc = acc + n[i] * row;
        }
        return acc;
    }

    // build_single_residual, voxel_mapping.cpp:247-318
    void build_single_residual(const PV& pv, const OctoTree* nd, int layer, double sigma_num, bool& ok, double& prob, Ptpl& out) const {
        const double radius_k = 3;
        const double* pw = pv.pw;
        if (nd->plane.is_plane) {
            const Plane& pl = nd->plane;
            const float dis_to_plane = (float)std::fabs(((pl.normal[0] * pw[0] + pl.normal[1] * pw[1]) + pl.normal[2] * pw[2]) + (double)pl.d);
            const float dis_to_center = (float)(((pl.center[0] - pw[0]) * (pl.center[0] - pw[0]) + (pl.center[1] - pw[1]) * (pl.center[1] - pw[1])) +
                                                (pl.center[2] - pw[2]) * (pl.center[2] - pw[2]));
            const float range_dis = std::sqrt(dis_to_center - dis_to_plane * dis_to_plane);
            if ((double)range_dis <= radius_k * (double)pl.radius) {
                double sigma_l = plane_sigma(pw, pl.center, pl.normal, pl.plane_var);
                sigma_l = sigma_l + quad_sym6(pl.normal, pv.var);
                const double sq = std::sqrt(sigma_l);
                if ((double)dis_to_plane < sigma_num * sq) {
                    ok = true;
                    const double dd = (double)dis_to_plane;
                    const double this_prob = 1.0 / sq * det_exp(-0.5 * dd * dd / sigma_l);
                    if (this_prob > prob) {
                        prob = this_prob;
                        for (int i = 0; i < 3; ++i) { out.point[i] = pv.pb[i]; out.normal[i] = pl.normal[i]; out.center[i] = pl.center[i]; }
                        for (int i = 0; i < 21; ++i) out.plane_var[i] = pl.plane_var[i];
                        out.d = (double)pl.d;
                        out.layer = layer;
                    }
                }
            }
            return;
        }
        if (layer < cfg.max_layer) {
            for (int l = 0; l < 8; ++l)
                if (nd->leaves[l] != nullptr) build_single_residual(pv, nd->leaves[l], layer + 1, sigma_num, ok, prob, out);
        }
    }

    // BuildResidualListOMP, voxel_mapping.cpp:153-245
    void build_residual_list(const std::vector<PV>& pv_list, double sigma_num, std::vector<Ptpl>& out) const {
        const int n = (int)pv_list.size();
        std::vector<Ptpl> all(n);
        std::vector<char> useful(n, 0);
        const double vs = cfg.voxel_size;  // double on the lookup path (:153)
#pragma omp parallel for num_threads(cfg.omp_threads > 0 ? cfg.omp_threads : 1) schedule(static)
        for (int i = 0; i < n; ++i) {
            const PV& pv = pv_list[i];
            float loc[3];
            int64_t kk[3];
            for (int j = 0; j < 3; ++j) {
                loc[j] = (float)(pv.pw[j] / vs);
                if (loc[j] < 0) loc[j] = (float)((double)loc[j] - 1.0);
                kk[j] = (int64_t)loc[j];
            }
            const VoxelKey key{kk[0], kk[1], kk[2]};
            auto it = feat_map.find(key);
            if (it == feat_map.end()) continue;
            const OctoTree* cur = it->second;
            Ptpl single;
            bool ok = false;
            double prob = 0;
            build_single_residual(pv, cur, 0, sigma_num, ok, prob, single);
            if (!ok) {
                VoxelKey nk = key;
                // unit mismatch (voxel units vs metres) replicated on purpose, :193-216
                const double ql = (double)cur->quater_length;
                if ((double)loc[0] > cur->voxel_center[0] + ql) nk.x = nk.x + 1;
                else if ((double)loc[0] < cur->voxel_center[0] - ql) nk.x = nk.x - 1;
                if ((double)loc[1] > cur->voxel_center[1] + ql) nk.y = nk.y + 1;
                else if ((double)loc[1] < cur->voxel_center[1] - ql) nk.y = nk.y - 1;
                if ((double)loc[2] > cur->voxel_center[2] + ql) nk.z = nk.z + 1;
                else if ((double)loc[2] < cur->voxel_center[2] - ql) nk.z = nk.z - 1;
                auto itn = feat_map.find(nk);
                if (itn != feat_map.end()) build_single_residual(pv, itn->second, 0, sigma_num, ok, prob, single);
            }
            if (ok) {
                single.src_index = i;
                all[i] = single;
                useful[i] = 1;
            }
        }
        out.clear();
        for (int i = 0; i < n; ++i)
            if (useful[i]) out.push_back(all[i]);
    }

    // per-scan body covariances + cross matrices, voxel_mapping.cpp:1302-1316
    void prepare_scan(const float* body, int n) {
        body_cov.assign((size_t)n * 6, 0.0);
        cross_mat.assign((size_t)n * 9, 0.0);
        for (int i = 0; i < n; ++i) {
            double p[3] = {(double)body[i * 3 + 0], (double)body[i * 3 + 1], (double)body[i * 3 + 2]};
            if (p[2] == 0) p[2] = 0.001;
            calc_body_var(p, dept_err_f, dir_var, &body_cov[(size_t)i * 6]);
            double q[3];
            mat3_vec(cfg.extR, p, q);
            for (int j = 0; j < 3; ++j) q[j] = q[j] + cfg.extT[j];
            skew(q, &cross_mat[(size_t)i * 9]);
        }
    }

    // world covariance used for matching, voxel_mapping.cpp:1356
    void world_cov_match(const double* R, const double* cov_state, int i, double* out6) const {
        double Sb[9], C[9], nC[9], rot_var[9], T1[6], T2[6];
        sym6_to_full(&body_cov[(size_t)i * 6], Sb);
        congr_sym6(R, Sb, T1);
        for (int k = 0; k < 9; ++k) { C[k] = cross_mat[(size_t)i * 9 + k]; nC[k] = -C[k]; }
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) rot_var[a * 3 + b] = cov_state[a * 18 + b];
        congr_sym6(nC, rot_var, T2);
        for (int a = 0; a < 3; ++a)
            for (int b = a; b < 3; ++b) out6[sym6_idx(a, b)] = (T1[sym6_idx(a, b)] + T2[sym6_idx(a, b)]) + cov_state[(3 + a) * 18 + (3 + b)];
    }

    // voxel_map_init, voxel_mapping.cpp:1243-1281 (uses the full undistorted scan)
    void voxel_map_init(const float* body, int n) {
        std::vector<PV> pv_list(n);
        for (int i = 0; i < n; ++i) {
            PV& pv = pv_list[i];
            const double pb[3] = {(double)body[i * 3 + 0], (double)body[i * 3 + 1], (double)body[i * 3 + 2]};
            double pw[3];
            body_to_world_f(state.rot, state.pos, pb, pw);
            for (int j = 0; j < 3; ++j) { pv.pb[j] = pw[j]; pv.pw[j] = pw[j]; }
            double pt[3] = {pb[0], pb[1], pb[2]};
            double bv[6], Sb[9], T1[6], T2[6], C[9], nC[9], rot_var[9];
            calc_body_var(pt, dept_err_f, dir_var, bv);
            skew(pt, C);  // cross matrix of the LiDAR-frame point (:1260-1261)
            for (int k = 0; k < 9; ++k) nC[k] = -C[k];
            sym6_to_full(bv, Sb);
            congr_sym6(state.rot, Sb, T1);
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) rot_var[a * 3 + b] = state.cov[a * 18 + b];
            congr_sym6(nC, rot_var, T2);
            for (int a = 0; a < 3; ++a)
                for (int b = a; b < 3; ++b) pv.var[sym6_idx(a, b)] = (T1[sym6_idx(a, b)] + T2[sym6_idx(a, b)]) + state.cov[(3 + a) * 18 + (3 + b)];
        }
        build_voxel_map(pv_list);
    }

    // state_propagat - state, common_lib.h:249-260
    static void state_minus(const State& a, const State& b, double* out) {
        double rotd[9];
        // b.rot^T * a.rot
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                rotd[i * 3 + j] = (b.rot[0 * 3 + i] * a.rot[0 * 3 + j] + b.rot[1 * 3 + i] * a.rot[1 * 3 + j]) + b.rot[2 * 3 + i] * a.rot[2 * 3 + j];
        so3_log(rotd, out);
        for (int i = 0; i < 3; ++i) {
            out[3 + i] = a.pos[i] - b.pos[i];
            out[6 + i] = a.vel[i] - b.vel[i];
            out[9 + i] = a.bg[i] - b.bg[i];
            out[12 + i] = a.ba[i] - b.ba[i];
            out[15 + i] = a.grav[i] - b.grav[i];
        }
    }
    static void state_plus(State& s, const double* add) {  // operator+=, common_lib.h:238-247
        double E[9], Rn[9];
        so3_exp(add[0], add[1], add[2], E);
        mat3_mul(s.rot, E, Rn);
        for (int i = 0; i < 9; ++i) s.rot[i] = Rn[i];
        for (int i = 0; i < 3; ++i) {
            s.pos[i] = s.pos[i] + add[3 + i];
            s.vel[i] = s.vel[i] + add[6 + i];
            s.bg[i] = s.bg[i] + add[9 + i];
            s.ba[i] = s.ba[i] + add[12 + i];
            s.grav[i] = s.grav[i] + add[15 + i];
        }
    }

    // lio_state_estimation, voxel_mapping.cpp:1284-1652.  `body` = down-sampled body-frame scan.
    void lio_state_estimation(const float* body, int n, const State& state_propagat) {
        iter_stats.clear();
        iters_run = 0;
        last_match_idx.clear();
        last_match_layer.clear();
        prepare_scan(body, n);
        if (feat_map.empty()) return;
        int rematch_num = 0;
        double G[324], HTH18[324];
        for (int i = 0; i < 324; ++i) { G[i] = 0; HTH18[i] = 0; }
        std::vector<PV> pv_list(n);
        std::vector<Ptpl> ptpl_list;
        for (int iter = 0; iter < cfg.max_iteration; ++iter) {
            // (a) world points + covariances, :1344-1359
            for (int i = 0; i < n; ++i) {
                PV& pv = pv_list[i];
                for (int j = 0; j < 3; ++j) pv.pb[j] = (double)body[i * 3 + j];
                body_to_world_f(state.rot, state.pos, pv.pb, pv.pw);
                world_cov_match(state.rot, state.cov, i, pv.var);
            }
            // (b) :1365 (sigma_num is the literal 3.0)
            build_residual_list(pv_list, 3.0, ptpl_list);
            const int m = (int)ptpl_list.size();
            // (c)+(d)+(e): rows, weights and normal equations, :1372-1392, :1487-1586
            IterStats st;
            FxAcc fx[27], fx_res;
            double sHTH[21], sHTz[6], sres = 0;
            for (double& v : sHTH) v = 0;
            for (double& v : sHTz) v = 0;
            double RRe[9];  // state.rot_end * m_extR
            mat3_mul(state.rot, cfg.extR, RRe);
            for (int i = 0; i < m; ++i) {
                const Ptpl& pt = ptpl_list[i];
                // float-rounded copies through PCL point structs (:1377-1389)
                const double pbf[3] = {(double)(float)pt.point[0], (double)(float)pt.point[1], (double)(float)pt.point[2]};
                const float nf[3] = {(float)pt.normal[0], (float)pt.normal[1], (float)pt.normal[2]};
                double pwd[3];
                body_to_world_d(state.rot, state.pos, pt.point, pwd);
                const float dis = (float)(((pwd[0] * (double)nf[0] + pwd[1] * (double)nf[1]) + pwd[2] * (double)nf[2]) + pt.d);
                // :1496-1527
                double p_imu[3];
                mat3_vec(cfg.extR, pbf, p_imu);
                for (int j = 0; j < 3; ++j) p_imu[j] = p_imu[j] + cfg.extT[j];
                double C[9];
                skew(p_imu, C);
                const double nv[3] = {(double)nf[0], (double)nf[1], (double)nf[2]};
                double pw2[3];
                mat3_vec(state.rot, p_imu, pw2);
                for (int j = 0; j < 3; ++j) pw2[j] = pw2[j] + state.pos[j];
                double bv[6], Sb[9], var6[6];
                double p_for_var[3] = {p_imu[0], p_imu[1], p_imu[2]};
                calc_body_var(p_for_var, dept_err_f, cfg.calib_laser ? dir_var_calib : dir_var, bv);
                sym6_to_full(bv, Sb);
                congr_sym6(RRe, Sb, var6);
                const double sigma_l = plane_sigma(pw2, pt.center, pt.normal, pt.plane_var);
                const double R_inv = 1.0 / (sigma_l + quad_sym6(nv, var6));
                // A = point_crossmat * rot^T * norm_vec  (:1562)
                double CRt[9], A[3];
                mat3_mul_bt(C, state.rot, CRt);  // (point_crossmat * rot^T) first, Eigen's left-to-right product
                mat3_vec(CRt, nv, A);
                const double h[6] = {A[0], A[1], A[2], nv[0], nv[1], nv[2]};
                const double z = -(double)dis;
                double hw[6];
                for (int a = 0; a < 6; ++a) hw[a] = h[a] * R_inv;  // Hsub_T_R_inv column (:1564)
                int e = 0;
                for (int a = 0; a < 6; ++a)
                    for (int b = a; b < 6; ++b, ++e) {
                        const double term = hw[a] * h[b];
                        fx[e].add(term);
                        sHTH[e] += term;
                    }
                for (int a = 0; a < 6; ++a) {
                    const double term = hw[a] * z;
                    fx[21 + a].add(term);
                    sHTz[a] += term;
                }
                fx_res.add(std::fabs((double)dis));
                sres += std::fabs((double)dis);
            }
            {
                int e = 0;
                for (int a = 0; a < 6; ++a)
                    for (int b = a; b < 6; ++b, ++e) {
                        const double v = cfg.sum_mode == 0 ? fx[e].value() : sHTH[e];
                        st.HTH[a * 6 + b] = v;
                        st.HTH[b * 6 + a] = v;
                    }
                for (int a = 0; a < 6; ++a) st.HTz[a] = cfg.sum_mode == 0 ? fx[21 + a].value() : sHTz[a];
                st.n_match = m;
                st.total_residual = cfg.sum_mode == 0 ? fx_res.value() : sres;
            }
            // (e) :1586-1592.  The reference forms K1 = (H^T R^-1 H (+) 0_12 + P^-1)^-1 with two 18x18 Eigen inverses.  H^T R^-1 H
            // is non-zero only in its 6x6 pose block A, so by the matrix inversion lemma
            //     K1[:, :6] = P[:, :6] (I6 + A P11)^-1,     P11 = P[:6, :6],
            // and K1 is only ever used through its first six columns (G = K1[:, :6] A, the solution, (I - G) P).  solve_mode 0
            // (default, what the CUDA product computes, bit for bit): that 6x6 form; solve_mode 1: the reference's literal two
            // inversions (Eigen's .inverse() is unpinned either way; tests/test_oracle_crosscheck.py bounds the gap between the two).
            double K1c[108];   // K1[:, :6], 18 x 6
            if (cfg.solve_mode == 0) {
                double B[36], S[36];
                for (int i = 0; i < 6; ++i)
                    for (int j = 0; j < 6; ++j) {
                        double s = (i == j) ? 1.0 : 0.0;
                        for (int k = 0; k < 6; ++k) s = s + st.HTH[i * 6 + k] * state.cov[k * 18 + j];
                        B[i * 6 + j] = s;
                    }
                lu_inverse<6>(B, S);
                for (int i = 0; i < 18; ++i)
                    for (int j = 0; j < 6; ++j) {
                        double s = 0.0;
                        for (int k = 0; k < 6; ++k) s = s + state.cov[i * 18 + k] * S[k * 6 + j];
                        K1c[i * 6 + j] = s;
                    }
            } else {
                for (int a = 0; a < 6; ++a)
                    for (int b = 0; b < 6; ++b) HTH18[a * 18 + b] = st.HTH[a * 6 + b];
                double Pinv[324], M[324], K1[324];
                lu_inverse<18>(state.cov, Pinv);
                for (int i = 0; i < 324; ++i) M[i] = HTH18[i] + Pinv[i];
                lu_inverse<18>(M, K1);
                for (int i = 0; i < 18; ++i)
                    for (int j = 0; j < 6; ++j) K1c[i * 6 + j] = K1[i * 18 + j];
            }
            for (int i = 0; i < 18; ++i)
                for (int j = 0; j < 6; ++j) {
                    double s = 0.0;
                    for (int k = 0; k < 6; ++k) s = s + K1c[i * 6 + k] * st.HTH[k * 6 + j];
                    G[i * 18 + j] = s;
                }
            double vec[18];
            state_minus(state_propagat, state, vec);
            double solution[18];
            for (int i = 0; i < 18; ++i) {
                double s1 = 0.0, s2 = 0.0;
                for (int k = 0; k < 6; ++k) s1 = s1 + K1c[i * 6 + k] * st.HTz[k];
                for (int k = 0; k < 6; ++k) s2 = s2 + G[i * 18 + k] * vec[k];
                solution[i] = (s1 + vec[i]) - s2;
            }
            state_plus(state, solution);
            for (int i = 0; i < 18; ++i) st.solution[i] = solution[i];
            // (f) :1619-1650
            const double rn = std::sqrt((solution[0] * solution[0] + solution[1] * solution[1]) + solution[2] * solution[2]);
            const double tn = std::sqrt((solution[3] * solution[3] + solution[4] * solution[4]) + solution[5] * solution[5]);
            const bool converged = (rn * 57.3 < 0.01) && (tn * 100 < 0.015);
            st.converged = converged ? 1 : 0;
            iter_stats.push_back(st);
            iters_run = iter + 1;
            last_match_idx.clear();
            last_match_layer.clear();
            for (const Ptpl& p : ptpl_list) { last_match_idx.push_back(p.src_index); last_match_layer.push_back(p.layer); }
            if (converged || ((rematch_num == 0) && (iter == cfg.max_iteration - 2))) rematch_num++;
            if (rematch_num >= 2 || iter == cfg.max_iteration - 1) {
                // cov = (I - G) * cov (:1646).  G has six non-zero columns: solve_mode 0 evaluates P - G[:, :6] P[:6, :]
                double ncov[324];
                for (int i = 0; i < 18; ++i)
                    for (int j = 0; j < 18; ++j) {
                        if (cfg.solve_mode == 0) {
                            double s = 0.0;
                            for (int k = 0; k < 6; ++k) s = s + G[i * 18 + k] * state.cov[k * 18 + j];
                            ncov[i * 18 + j] = state.cov[i * 18 + j] - s;
                        } else {
                            double s = 0.0;
                            for (int k = 0; k < 18; ++k) {
                                const double ig = ((i == k) ? 1.0 : 0.0) - G[i * 18 + k];
                                s = s + ig * state.cov[k * 18 + j];
                            }
                            ncov[i * 18 + j] = s;
                        }
                    }
                for (int i = 0; i < 324; ++i) state.cov[i] = ncov[i];
                break;
            }
        }
    }

    // map_incremental_grow (VoxelMap part), ImMesh_mesh_reconstruction.cpp:387-408
    void map_incremental_grow(const float* body, int n) {
        std::vector<PV> pv_list(n);
        double RRe[9];
        mat3_mul(state.rot, cfg.extR, RRe);
        for (int i = 0; i < n; ++i) {
            PV& pv = pv_list[i];
            const double pb[3] = {(double)body[i * 3 + 0], (double)body[i * 3 + 1], (double)body[i * 3 + 2]};
            double pw[3];
            body_to_world_f(state.rot, state.pos, pb, pw);
            for (int j = 0; j < 3; ++j) { pv.pb[j] = pw[j]; pv.pw[j] = pw[j]; }
            double Sb[9], T1[6], T2[6], C[9], nC[9], rot_var[9];
            sym6_to_full(&body_cov[(size_t)i * 6], Sb);
            congr_sym6(RRe, Sb, T1);
            for (int k = 0; k < 9; ++k) { C[k] = cross_mat[(size_t)i * 9 + k]; nC[k] = -C[k]; }
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) rot_var[a * 3 + b] = state.cov[a * 18 + b];
            congr_sym6(nC, rot_var, T2);
            for (int a = 0; a < 3; ++a)
                for (int b = a; b < 3; ++b) pv.var[sym6_idx(a, b)] = (T1[sym6_idx(a, b)] + T2[sym6_idx(a, b)]) + state.cov[(3 + a) * 18 + (3 + b)];
        }
        // std::sort(pv_list, var_contrast): ascending ||diag(var)|| (voxel_mapping.cpp:49). Ties are broken by
        // scan index here (libstdc++ introsort's tie order is unspecified; ties do not occur on noisy data).
        std::vector<std::pair<double, int>> keyed(n);
        for (int i = 0; i < n; ++i) {
            const double* v = pv_list[i].var;
            keyed[i] = {std::sqrt((v[0] * v[0] + v[3] * v[3]) + v[5] * v[5]), i};
        }
        std::sort(keyed.begin(), keyed.end());
        std::vector<PV> sorted(n);
        for (int i = 0; i < n; ++i) sorted[i] = pv_list[keyed[i].second];
        update_voxel_map(sorted);
    }

    // Forward_without_imu (constant-velocity prediction), IMU_Processing.cpp:486-553.  cov_acc = cov_gyr = 0.1
    // (IMU_Processing.cpp:56-57) unless configured.
    void forward_without_imu(double dt, double cov_gyr, double cov_acc) {
        double Expf[9], Fx[324], cw[324];
        so3_exp_dt(state.bg, dt, Expf);
        for (int i = 0; i < 324; ++i) { Fx[i] = 0; cw[i] = 0; }
        for (int i = 0; i < 18; ++i) Fx[i * 18 + i] = 1.0;
        double En[9];
        so3_exp_dt(state.bg, -dt, En);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) Fx[a * 18 + b] = En[a * 3 + b];
        for (int a = 0; a < 3; ++a) { Fx[a * 18 + 9 + a] = dt; Fx[(3 + a) * 18 + 6 + a] = dt; }
        for (int a = 0; a < 3; ++a) { cw[(9 + a) * 18 + 9 + a] = cov_gyr * dt * dt; cw[(6 + a) * 18 + 6 + a] = cov_acc * dt * dt; }
        double T[324], ncov[324];
        for (int i = 0; i < 18; ++i)
            for (int j = 0; j < 18; ++j) {
                double s = 0.0;
                for (int k = 0; k < 18; ++k) s = s + Fx[i * 18 + k] * state.cov[k * 18 + j];
                T[i * 18 + j] = s;
            }
        for (int i = 0; i < 18; ++i)
            for (int j = 0; j < 18; ++j) {
                double s = 0.0;
                for (int k = 0; k < 18; ++k) s = s + T[i * 18 + k] * Fx[j * 18 + k];
                ncov[i * 18 + j] = s + cw[i * 18 + j];
            }
        for (int i = 0; i < 324; ++i) state.cov[i] = ncov[i];
        double Rn[9];
        mat3_mul(state.rot, Expf, Rn);
        for (int i = 0; i < 9; ++i) state.rot[i] = Rn[i];
        for (int i = 0; i < 3; ++i) state.pos[i] = state.pos[i] + state.vel[i] * dt;
    }

    // ------------------------------------------------------------------ canonical map dump (for parity tests)
    // One row of 44 doubles per octree node, roots sorted by key, nodes in pre-order (child 0..7):
    // [kx,ky,kz, path, layer, init_octo, is_plane, update_enable, n_temp, new_points, octo_state(unused=0),
    //  cx,cy,cz, nx,ny,nz, d, radius, min_eig, points_size, vcx,vcy,vcz(voxel centre) , plane_var[21]] -> 3+8+3+3+4+3+21 = 45
    static const int kDumpCols = 45;
    void dump_node(const VoxelKey& k, const OctoTree* nd, int64_t path, std::vector<double>& out) const {
        double row[kDumpCols];
        int c = 0;
        row[c++] = (double)k.x; row[c++] = (double)k.y; row[c++] = (double)k.z;
        row[c++] = (double)path; row[c++] = nd->layer; row[c++] = nd->init_octo; row[c++] = nd->plane.is_plane;
        row[c++] = nd->update_enable; row[c++] = (double)nd->temp_points.size(); row[c++] = nd->new_points; row[c++] = 0;
        const bool pl = nd->plane.is_plane;
        for (int i = 0; i < 3; ++i) row[c++] = nd->plane.center[i];
        for (int i = 0; i < 3; ++i) row[c++] = pl ? nd->plane.normal[i] : 0.0;
        row[c++] = pl ? nd->plane.d : 0.0; row[c++] = pl ? nd->plane.radius : 0.0; row[c++] = pl ? nd->plane.min_eigen_value : 0.0;
        row[c++] = nd->plane.points_size;
        for (int i = 0; i < 3; ++i) row[c++] = nd->voxel_center[i];
        for (int i = 0; i < 21; ++i) row[c++] = pl ? nd->plane.plane_var[i] : 0.0;
        out.insert(out.end(), row, row + kDumpCols);
        for (int l = 0; l < 8; ++l)
            if (nd->leaves[l]) dump_node(k, nd->leaves[l], path * 9 + (l + 1), out);
    }
    void dump_map(std::vector<double>& out) const {
        std::vector<VoxelKey> keys;
        for (auto& kv : feat_map) keys.push_back(kv.first);
        std::sort(keys.begin(), keys.end(), [](const VoxelKey& a, const VoxelKey& b) {
            if (a.x != b.x) return a.x < b.x;
            if (a.y != b.y) return a.y < b.y;
            return a.z < b.z;
        });
        out.clear();
        for (const VoxelKey& k : keys) dump_node(k, feat_map.at(k), 0, out);
    }
};

}  // namespace orc
