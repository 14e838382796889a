// ORACLE (test infrastructure, NOT product code).
// CPU restatement helpers for the ImMesh hot path.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use anything under oracle/.
//
// Parity status: "parity unpinned" -- the reference ships no tests / golden vectors and
// cannot be compiled here (needs ROS, Eigen, PCL, CGAL, TBB).  This
// restatement follows the reference sources line by line (citations per function) and
// is cross-checked against numpy / scipy (tests/test_oracle_crosscheck.py).
//
// Numerical conventions (shared, by specification, with the CUDA product so that every
// discrete decision -- voxel keys, octree shape, match sets, facets -- is bit-reproducible):
//   * IEEE-754 binary64 / binary32, round-to-nearest-even, NO fused multiply-add
//     (build with -ffp-contract=off; the reference builds with -O3 -msse2 and no -mfma,
//     CMakeLists.txt:14, so it has no contraction either).
//   * sums are evaluated left-to-right in the order written here.
//   * sin/cos/exp/acos are the arithmetic-only implementations below (libm results differ
//     between glibc and CUDA in the last ulp; these differ from glibc by <= a few ulp).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

// ---------------------------------------------------------------- deterministic libm subset
static const double kPio2Hi = 1.57079632673412561417e+00;  // 33 bits of pi/2
static const double kPio2Lo = 6.07710050650619224932e-11;  // pi/2 - kPio2Hi
static const double kTwoOverPi = 6.36619772367581382433e-01;
static const double kLn2Hi = 6.93147180369123816490e-01;
static const double kLn2Lo = 1.90821492927058770002e-10;
static const double kInvLn2 = 1.44269504088896338700e+00;
static const double kPi = 3.14159265358979311600e+00;

// Taylor kernels on |r| <= pi/4 (Horner in r^2, fixed coefficient order)
inline double sin_kernel(double r) {
    const double z = r * r;
    double p = -1.0 / 51090942171709440000.0;  // -1/21!
    p = p * z + 1.0 / 121645100408832000.0;    // 1/19!
    p = p * z - 1.0 / 355687428096000.0;       // 1/17!
    p = p * z + 1.0 / 1307674368000.0;         // 1/15!
    p = p * z - 1.0 / 6227020800.0;            // 1/13!
    p = p * z + 1.0 / 39916800.0;              // 1/11!
    p = p * z - 1.0 / 362880.0;                // 1/9!
    p = p * z + 1.0 / 5040.0;                  // 1/7!
    p = p * z - 1.0 / 120.0;                   // 1/5!
    p = p * z + 1.0 / 6.0;                     // 1/3!  (sign folded below)
    return r - (r * z) * p;
}
inline double cos_kernel(double r) {
    const double z = r * r;
    double p = 1.0 / 2432902008176640000.0;  // 1/20!
    p = p * z - 1.0 / 6402373705728000.0;    // 1/18!
    p = p * z + 1.0 / 20922789888000.0;      // 1/16!
    p = p * z - 1.0 / 87178291200.0;         // 1/14!
    p = p * z + 1.0 / 479001600.0;           // 1/12!
    p = p * z - 1.0 / 3628800.0;             // 1/10!
    p = p * z + 1.0 / 40320.0;               // 1/8!
    p = p * z - 1.0 / 720.0;                 // 1/6!
    p = p * z + 1.0 / 24.0;                  // 1/4!
    p = p * z - 1.0 / 2.0;                   // 1/2!
    return 1.0 + z * p;
}
inline void det_sincos(double x, double* s, double* c) {
    const double kf = std::nearbyint(x * kTwoOverPi);
    const double r = (x - kf * kPio2Hi) - kf * kPio2Lo;
    const long long k = (long long)kf;
    const double sk = sin_kernel(r), ck = cos_kernel(r);
    switch ((int)(k & 3)) {
        case 0: *s = sk; *c = ck; break;
        case 1: *s = ck; *c = -sk; break;
        case 2: *s = -sk; *c = -ck; break;
        default: *s = -ck; *c = sk; break;
    }
}
inline double det_sin(double x) { double s, c; det_sincos(x, &s, &c); return s; }
inline double det_cos(double x) { double s, c; det_sincos(x, &s, &c); return c; }

// exp(x) for x <= 0 (used for the match probability, voxel_mapping.cpp:272)
inline double det_exp(double x) {
    if (!(x > -700.0)) return 0.0;
    const double kf = std::nearbyint(x * kInvLn2);
    const double r = (x - kf * kLn2Hi) - kf * kLn2Lo;
    double p = 1.0 / 87178291200.0;  // 1/14!
    p = p * r + 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    return std::ldexp(p, (int)kf);
}

// asin on |s| <= 0.5 by its Maclaurin series with a run-time coefficient recurrence
inline double asin_small(double s) {
    const double z = s * s;
    double term = s, sum = s, c = 1.0;
    for (int k = 0; k < 40; ++k) {
        const double a = (double)(2 * k + 1);
        c = c * (a * a) / ((double)(2 * k + 2) * (double)(2 * k + 3));
        term = term * z;
        const double add = c * term;
        sum = sum + add;
        if (std::fabs(add) < 1e-19 * std::fabs(sum)) break;
    }
    return sum;
}
inline double det_acos(double x) {
    if (x >= 1.0) return 0.0;
    if (x <= -1.0) return kPi;
    if (x > 0.5) return 2.0 * asin_small(std::sqrt((1.0 - x) * 0.5));
    if (x < -0.5) return kPi - 2.0 * asin_small(std::sqrt((1.0 + x) * 0.5));
    return (kPio2Hi - asin_small(x)) + kPio2Lo;
}

inline double det_asin(double x) {
    const double ax = std::fabs(x);
    double r;
    if (ax <= 0.5) r = asin_small(ax);
    else r = (kPio2Hi - 2.0 * asin_small(std::sqrt((1.0 - ax) * 0.5))) + kPio2Lo;
    return x < 0 ? -r : r;
}
inline double det_atan2(double y, double x) {
    if (x == 0.0 && y == 0.0) return 0.0;
    const double r = std::sqrt(x * x + y * y);
    const double pio2 = 1.57079632679489655800e+00;
    if (std::fabs(x) >= std::fabs(y)) {
        const double a = det_asin(y / r);
        if (x > 0) return a;
        return (y >= 0 ? kPi : -kPi) - a;
    }
    const double a = det_asin(x / r);
    return y > 0 ? (pio2 - a) : (a - pio2);
}

// ---------------------------------------------------------------- 3x3 helpers (row-major)
inline void mat3_mul(const double* A, const double* B, double* C) {  // C = A*B
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = (A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j]) + A[i * 3 + 2] * B[2 * 3 + j];
}
inline void mat3_mul_bt(const double* A, const double* B, double* C) {  // C = A*B^T
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = (A[i * 3 + 0] * B[j * 3 + 0] + A[i * 3 + 1] * B[j * 3 + 1]) + A[i * 3 + 2] * B[j * 3 + 2];
}
inline void mat3_vec(const double* A, const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = (A[i * 3 + 0] * v[0] + A[i * 3 + 1] * v[1]) + A[i * 3 + 2] * v[2];
}
inline void mat3_tvec(const double* A, const double* v, double* o) {  // A^T v
    for (int i = 0; i < 3; ++i) o[i] = (A[0 * 3 + i] * v[0] + A[1 * 3 + i] * v[1]) + A[2 * 3 + i] * v[2];
}
inline void skew(const double* v, double* K) {  // SKEW_SYM_MATRX, so3_math.h:9
    K[0] = 0.0; K[1] = -v[2]; K[2] = v[1];
    K[3] = v[2]; K[4] = 0.0; K[5] = -v[0];
    K[6] = -v[1]; K[7] = v[0]; K[8] = 0.0;
}
inline double dot3(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline double norm3(const double* a) { return std::sqrt(dot3(a, a)); }

// symmetric 3x3 stored as 6: [00,01,02,11,12,22]
inline int sym6_idx(int i, int j) {
    static const int t[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
    return t[i * 3 + j];
}
inline void sym6_to_full(const double* s, double* M) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i * 3 + j] = s[sym6_idx(i, j)];
}
// S = A * B * A^T for symmetric B (full 3x3 in), upper triangle evaluated, returned as sym6
inline void congr_sym6(const double* A, const double* Bfull, double* out6) {
    double T[9];
    mat3_mul(A, Bfull, T);
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j)
            out6[sym6_idx(i, j)] = (T[i * 3 + 0] * A[j * 3 + 0] + T[i * 3 + 1] * A[j * 3 + 1]) + T[i * 3 + 2] * A[j * 3 + 2];
}

// SO(3) Exp(v1,v2,v3), so3_math.h:54-72 (threshold 1e-5)
inline void so3_exp(double v1, double v2, double v3, double* R) {
    const double n = std::sqrt((v1 * v1 + v2 * v2) + v3 * v3);
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    if (n > 0.00001) {
        const double r[3] = {v1 / n, v2 / n, v3 / n};
        double K[9], KK[9];
        skew(r, K);
        mat3_mul(K, K, KK);
        double s, c;
        det_sincos(n, &s, &c);
        const double omc = 1.0 - c;
        for (int i = 0; i < 9; ++i) R[i] = (R[i] + s * K[i]) + omc * KK[i];
    }
}
// SO(3) Exp(ang_vel, dt), so3_math.h:31-51 (threshold 1e-7)
inline void so3_exp_dt(const double* w, double dt, double* R) {
    const double n = norm3(w);
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    if (n > 0.0000001) {
        const double r[3] = {w[0] / n, w[1] / n, w[2] / n};
        double K[9], KK[9];
        skew(r, K);
        mat3_mul(K, K, KK);
        double s, c;
        det_sincos(n * dt, &s, &c);
        const double omc = 1.0 - c;
        for (int i = 0; i < 9; ++i) R[i] = (R[i] + s * K[i]) + omc * KK[i];
    }
}
// SO(3) Log, so3_math.h:75-81
inline void so3_log(const double* R, double* out) {
    const double tr = (R[0] + R[4]) + R[8];
    const double theta = (tr > 3.0 - 1e-6) ? 0.0 : det_acos(0.5 * (tr - 1.0));
    const double K[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    if (std::fabs(theta) < 0.001) {
        for (int i = 0; i < 3; ++i) out[i] = 0.5 * K[i];
    } else {
        const double f = 0.5 * theta / det_sin(theta);
        for (int i = 0; i < 3; ++i) out[i] = f * K[i];
    }
}

// ---------------------------------------------------------------- symmetric 3x3 eigen-solver
// Cyclic Jacobi (the reference calls Eigen::EigenSolver, voxel_loc.cpp:62, and
// Eigen::SelfAdjointEigenSolver, mesh_rec_geometry.cpp:199 -- Eigen is not vendored, so this
// restates the mathematical result; eigenvector sign is library-defined and every consumer
// on the path is sign-invariant).  a: sym6 in, d: eigenvalues, V: eigenvectors in columns
// (row-major 3x3), unsorted.
inline void jacobi_eig3(const double* a6, double* d, double* V) {
    double a[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[i][j] = a6[sym6_idx(i, j)];
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = (std::fabs(a[0][1]) + std::fabs(a[0][2])) + std::fabs(a[1][2]);
        if (off == 0.0) break;
        for (int p = 0; p < 2; ++p) {
            for (int q = p + 1; q < 3; ++q) {
                const double apq = a[p][q];
                if (apq == 0.0) continue;
                const double g = 100.0 * std::fabs(apq);
                if (sweep > 3 && (std::fabs(a[p][p]) + g == std::fabs(a[p][p])) &&
                    (std::fabs(a[q][q]) + g == std::fabs(a[q][q]))) {
                    a[p][q] = 0.0;
                    a[q][p] = 0.0;
                    continue;
                }
                const double h = a[q][q] - a[p][p];
                double t;
                if (std::fabs(h) + g == std::fabs(h)) {
                    t = apq / h;
                } else {
                    const double theta = 0.5 * h / apq;
                    t = 1.0 / (std::fabs(theta) + std::sqrt(1.0 + theta * theta));
                    if (theta < 0.0) t = -t;
                }
                const double c = 1.0 / std::sqrt(1.0 + t * t);
                const double s = t * c;
                const double tau = s / (1.0 + c);
                const double hh = t * apq;
                a[p][p] = a[p][p] - hh;
                a[q][q] = a[q][q] + hh;
                a[p][q] = 0.0;
                a[q][p] = 0.0;
                const int r = 3 - p - q;  // the remaining index
                {
                    const double arp = a[r][p], arq = a[r][q];
                    const double nrp = arp - s * (arq + arp * tau);
                    const double nrq = arq + s * (arp - arq * tau);
                    a[r][p] = nrp; a[p][r] = nrp;
                    a[r][q] = nrq; a[q][r] = nrq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = vkp - s * (vkq + vkp * tau);
                    v[k][q] = vkq + s * (vkp - vkq * tau);
                }
            }
        }
    }
    for (int i = 0; i < 3; ++i) {
        d[i] = a[i][i];
        for (int j = 0; j < 3; ++j) V[i * 3 + j] = v[i][j];
    }
}

// ---------------------------------------------------------------- NxN inverse, LU partial pivoting
// (Eigen MatrixBase::inverse() for 18x18 = PartialPivLU, voxel_mapping.cpp:1588)
template <int N>
inline bool lu_inverse(const double* Ain, double* Ainv) {
    double a[N][N];
    int piv[N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) a[i][j] = Ain[i * N + j];
    for (int i = 0; i < N; ++i) piv[i] = i;
    bool ok = true;
    for (int k = 0; k < N; ++k) {
        int best = k;
        double bv = std::fabs(a[k][k]);
        for (int i = k + 1; i < N; ++i) {
            const double v = std::fabs(a[i][k]);
            if (v > bv) { bv = v; best = i; }
        }
        if (bv == 0.0) ok = false;
        if (best != k) {
            for (int j = 0; j < N; ++j) { const double t = a[k][j]; a[k][j] = a[best][j]; a[best][j] = t; }
            const int t = piv[k]; piv[k] = piv[best]; piv[best] = t;
        }
        const double pivv = a[k][k];
        for (int i = k + 1; i < N; ++i) a[i][k] = a[i][k] / pivv;
        for (int i = k + 1; i < N; ++i) {
            const double lik = a[i][k];
            for (int j = k + 1; j < N; ++j) a[i][j] = a[i][j] - lik * a[k][j];
        }
    }
    // solve A x = e_c for every column c
    for (int c = 0; c < N; ++c) {
        double y[N];
        for (int i = 0; i < N; ++i) {
            double s = (piv[i] == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; ++j) s = s - a[i][j] * y[j];
            y[i] = s;
        }
        for (int i = N - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < N; ++j) s = s - a[i][j] * y[j];
            y[i] = s / a[i][i];
        }
        for (int i = 0; i < N; ++i) Ainv[i * N + c] = y[i];
    }
    return ok;
}

// ---------------------------------------------------------------- order-free fixed-point accumulation
// Normal-equation terms are summed as 2^-20-quantised integers (hi/lo 32-bit split) so that
// the result does not depend on summation order (CPU serial loop, GPU atomics, any number of
// ranks).  The reference sums doubles inside an Eigen GEMM whose order is unspecified
// (voxel_mapping.cpp:1585-1586); the quantisation error is < 2^-21 per term.
static const double kFxScale = 1048576.0;  // 2^20
struct FxAcc {
    long long hi = 0, lo = 0;
    inline void add(double x) {
        const long long t = std::llrint(x * kFxScale);
        hi += (t >> 32);
        lo += (t & 0xffffffffLL);
    }
    inline double value() const { return ((double)hi * 4294967296.0 + (double)lo) * (1.0 / kFxScale); }
};

}  // namespace orc
