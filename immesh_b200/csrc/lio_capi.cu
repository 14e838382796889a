// immesh_b200 -- CUDA kernels (sm_100a) and C-ABI host orchestration of the localization path.
// Kernel bodies live in lio_core.cuh / voxelmap.cuh; this file adds the launch geometry, the
// block-level integer reduction of the normal equations, stream/event plumbing and the C ABI.
// There is no CPU path: every entry point needs a CUDA device and fails loudly without one.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/immesh_b200.h"
#include "common_host.hpp"
#include "peer_win.cuh"
#include "handles.hpp"
#include "map_dump.hpp"

using namespace immesh;

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(128) k_prepare(LioParams P, ScanBuf sb, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) prepare_point(P, sb, i);
}

__global__ void k_reset_scan(ScanBuf sb, LioCtrl* ctrl, int copy_prop) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nt = gridDim.x * blockDim.x;
    unsigned long long* acc = &ctrl->acc[0][0];
    for (int i = tid; i < IM_MAX_ITER * IM_NTERMS * 2; i += nt) acc[i] = 0ull;
    if (copy_prop)
        for (int i = tid; i < IM_STATE_DOUBLES; i += nt) ctrl->state_prop[i] = ctrl->state[i];
    if (tid == 0) {
        ctrl->stop = 0;
        ctrl->iters_run = 0;
        ctrl->rematch_num = 0;
        for (int i = 0; i < IM_MAX_ITER; ++i) ctrl->blocks_done[i] = 0;
    }
}

// ---- IESKF solve on the device, executed by the LAST residual block of an iteration (no extra launch).
// 18x18 partial-pivot LU inverse by one warp: lane i owns row i in registers, pivot row broadcast by shuffles.
// Element update order is exactly that of the serial algorithm (lu_inverse18 / the oracle): bit-identical results.
__device__ __noinline__ void warp_lu_inverse18(const double* a_in, double* lu /*[18][19] shared*/, int* piv_s, double* inv_out, int lane) {
    // rows live in shared memory with a padded stride (19) so that a column access by 18 lanes is nearly conflict free
    for (int idx = lane; idx < 324; idx += 32) lu[(idx / 18) * 19 + (idx % 18)] = a_in[idx];
    if (lane < 18) piv_s[lane] = lane;
    __syncwarp();
    for (int k = 0; k < 18; ++k) {
        // first maximum of |a[i][k]|, i >= k
        double bv = (lane >= k && lane < 18) ? fabs(lu[lane * 19 + k]) : -1.0;
        int bi = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (bi != k) {
            if (lane < 18) { const double t = lu[k * 19 + lane]; lu[k * 19 + lane] = lu[bi * 19 + lane]; lu[bi * 19 + lane] = t; }
            if (lane == 0) { const int t = piv_s[k]; piv_s[k] = piv_s[bi]; piv_s[bi] = t; }
        }
        __syncwarp();
        const double pivv = lu[k * 19 + k];
        __syncwarp();
        if (lane > k && lane < 18) lu[lane * 19 + k] = lu[lane * 19 + k] / pivv;
        __syncwarp();
        const int m = 17 - k;
        for (int idx = lane; idx < m * m; idx += 32) {
            const int i = k + 1 + idx / m, j = k + 1 + idx % m;
            lu[i * 19 + j] = lu[i * 19 + j] - lu[i * 19 + k] * lu[k * 19 + j];
        }
        __syncwarp();
    }
    if (lane < 18) {
        // column c of the inverse, substituted in place in inv_out (y[i] lives at inv_out[i][c])
        const int c = lane;
        for (int i = 0; i < 18; ++i) {
            double s = (piv_s[i] == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; ++j) s = s - lu[i * 19 + j] * inv_out[j * 18 + c];
            inv_out[i * 18 + c] = s;
        }
        for (int i = 17; i >= 0; --i) {
            double s = inv_out[i * 18 + c];
            for (int j = i + 1; j < 18; ++j) s = s - lu[i * 19 + j] * inv_out[j * 18 + c];
            inv_out[i * 18 + c] = s / lu[i * 19 + i];
        }
    }
    __syncwarp();
}

// ---- 18x18 inverse, block-parallel, bit-identical to the serial partial-pivot LU + substitution (lu_inverse18 / the oracle).
// Augmented elimination [A | I]: the forward substitution is folded into the LU -- element (i, c) of the right half receives
//   s = s - l_ik * y_k   for k = 0, 1, ...  exactly in the order (and with the operands) of  y_i = b_i - sum_{j<i} L_ij y_j,
// because a row's multipliers travel with it through the row swaps.  One thread per element of the 18 x 36 panel, one
// __syncthreads per pivot step: ping-pong buffers make the row swap a re-indexed read (row k of the step = old row `bi`),
// frozen rows go to U / Y.  The pivot is found redundantly by every thread (adjacent-pair tournament, left wins ties = first
// maximum), so no broadcast is needed.  The back substitution keeps the serial order (j ascending) per column: 18 lanes,
// registers only.  ~18 x (scan + div + update + barrier) + 18 x (div + short add chain) instead of one warp's 9 k
// dependent instructions.
#define INV_THREADS 672
struct InvScratch {
    double buf[2][18 * 37];
    double U[18 * 19];
    double Y[18 * 19];
};
__device__ __forceinline__ int pivot_row18(const double* m /*stride 37*/, int k) {
    double v[18];
    int id[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) { v[i] = (i >= k) ? fabs(m[i * 37 + k]) : -1.0; id[i] = i; }
#define IM_PMERGE(a, b) if (v[b] > v[a]) { v[a] = v[b]; id[a] = id[b]; }
    IM_PMERGE(0, 1) IM_PMERGE(2, 3) IM_PMERGE(4, 5) IM_PMERGE(6, 7) IM_PMERGE(8, 9) IM_PMERGE(10, 11) IM_PMERGE(12, 13) IM_PMERGE(14, 15) IM_PMERGE(16, 17)
    IM_PMERGE(0, 2) IM_PMERGE(4, 6) IM_PMERGE(8, 10) IM_PMERGE(12, 14)
    IM_PMERGE(0, 4) IM_PMERGE(8, 12)
    IM_PMERGE(0, 8)
    IM_PMERGE(0, 16)
#undef IM_PMERGE
    return id[0];
}
// variant 1 ("lean"): no redundant double-precision work -- the pivot is found by warp 0 alone (shuffle tournament, lower lane wins
// ties), the 17-k multipliers l_i are divided once per row, then every thread updates its element; three barriers per step.
__device__ long long g_inv_stamps[16];   // diagnostic: clock64 at the phase boundaries of the last block_lu_inverse18_lean call
__device__ __noinline__ void block_lu_inverse18_lean(const double* a_in, InvScratch* W, double* inv_out, bool rolled, bool redux) {
    __shared__ int s_bi;
    if (threadIdx.x == 0) g_inv_stamps[0] = clock64();
    __shared__ double s_l[18];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    for (int e = tid; e < 648; e += nt) {
        const int i = e / 36, j = e - i * 36;
        W->buf[0][i * 37 + j] = (j < 18) ? a_in[i * 18 + j] : ((j - 18 == i) ? 1.0 : 0.0);
    }
    __syncthreads();
    if (threadIdx.x == 0) g_inv_stamps[1] = clock64();
    int p = 0;
    for (int k = 0; k < 18; ++k) {
        const double* old = W->buf[p];
        double* nw = W->buf[p ^ 1];
        const bool stamp = (k == 5 && tid == 0);
        if (stamp) g_inv_stamps[8] = clock64();
        if (tid < 32 && redux) {
            // measured (clock64 stamps, profiles/README.md): the shuffle tournament below is a ~105-instruction dependent chain,
            // ~1.5 k cycles per step -- most of the kernel.  Non-negative doubles order like their bit patterns, so the first
            // maximum is three warp REDUX operations: max of the high words, max of the low words among the lanes that hold
            // that high word, min lane among the lanes that hold both.
            const bool act = lane >= k && lane < 18;
            const unsigned long long bits = act ? (unsigned long long)__double_as_longlong(fabs(old[lane * 37 + k])) : 0ull;
            const unsigned int hi = (unsigned int)(bits >> 32), lo = (unsigned int)bits;
            const unsigned int mhi = __reduce_max_sync(0xffffffffu, act ? hi : 0u);
            const bool c1 = act && hi == mhi;
            const unsigned int mlo = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
            const bool c2 = c1 && lo == mlo;
            const unsigned int bl = __reduce_min_sync(0xffffffffu, c2 ? (unsigned int)lane : 0xffffffffu);
            if (lane == 0) s_bi = (int)bl;
        } else if (tid < 32) {   // first maximum of |old[i][k]|, i >= k
            double bv = (lane >= k && lane < 18) ? fabs(old[lane * 37 + k]) : -1.0;
            int bi = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) s_bi = bi;
        }
        if (stamp) g_inv_stamps[9] = clock64();
        __syncthreads();
        if (stamp) g_inv_stamps[10] = clock64();
        const int bi = s_bi;
        if (tid > k && tid < 18) {   // multiplier of (post-swap) row tid
            const int src = (tid == bi) ? k : tid;
            s_l[tid] = old[src * 37 + k] / old[bi * 37 + k];
        }
        if (stamp) g_inv_stamps[11] = clock64();
        __syncthreads();
        if (stamp) g_inv_stamps[12] = clock64();
        for (int e = tid; e < 648; e += nt) {
            const int i = e / 36, j = e - i * 36;
            if (i < k || j < k) continue;
            if (i == k) {
                const double v = old[bi * 37 + j];
                if (j < 18) W->U[k * 19 + j] = v; else W->Y[k * 19 + (j - 18)] = v;
            } else if (j > k) {
                const int src = (i == bi) ? k : i;
                nw[i * 37 + j] = old[src * 37 + j] - s_l[i] * old[bi * 37 + j];
            }
        }
        if (stamp) g_inv_stamps[13] = clock64();
        __syncthreads();
        if (stamp) g_inv_stamps[14] = clock64();
        p ^= 1;
        if (k == 0 && threadIdx.x == 0) g_inv_stamps[2] = clock64();
    }
    if (threadIdx.x == 0) g_inv_stamps[3] = clock64();
    if (rolled) {
        // compact code (a straight-line unrolled substitution is ~900 instructions executed once: instruction fetch, not
        // arithmetic, then sets the pace).  x lives in inv_out (column c is private to thread c), j ascending as in the oracle.
        if (tid < 18) {
            const int c = tid;
#pragma unroll 1
            for (int i = 17; i >= 0; --i) {
                double s = W->Y[i * 19 + c];
#pragma unroll 2
                for (int j = i + 1; j < 18; ++j) s = s - W->U[i * 19 + j] * inv_out[j * 18 + c];
                inv_out[i * 18 + c] = s / W->U[i * 19 + i];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) g_inv_stamps[4] = clock64();
        return;
    }
    if (tid < 18) {
        const int c = tid;
        double x[18];
#pragma unroll
        for (int i = 17; i >= 0; --i) {
            double s = W->Y[i * 19 + c];
#pragma unroll
            for (int j = i + 1; j < 18; ++j) s = s - W->U[i * 19 + j] * x[j];
            x[i] = s / W->U[i * 19 + i];
        }
#pragma unroll
        for (int i = 0; i < 18; ++i) inv_out[i * 18 + c] = x[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) g_inv_stamps[4] = clock64();
}
__device__ __noinline__ void block_lu_inverse18_redundant(const double* a_in /*[324] shared, row-major*/, InvScratch* W, double* inv_out /*[324] shared*/) {

    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < 648; e += nt) {
        const int i = e / 36, j = e - i * 36;
        W->buf[0][i * 37 + j] = (j < 18) ? a_in[i * 18 + j] : ((j - 18 == i) ? 1.0 : 0.0);
    }
    __syncthreads();
    int p = 0;
    for (int k = 0; k < 18; ++k) {
        const double* old = W->buf[p];
        double* nw = W->buf[p ^ 1];
        const int bi = pivot_row18(old, k);
        const double pivv = old[bi * 37 + k];
        for (int e = tid; e < 648; e += nt) {
            const int i = e / 36, j = e - i * 36;
            if (i < k || j < k) continue;            // frozen rows, eliminated columns
            if (i == k) {                            // row k of this step (= old row bi) is final
                const double v = old[bi * 37 + j];
                if (j < 18) W->U[k * 19 + j] = v; else W->Y[k * 19 + (j - 18)] = v;
            } else if (j > k) {
                const int src = (i == bi) ? k : i;   // the swap, as a re-indexed read
                const double l = old[src * 37 + k] / pivv;
                nw[i * 37 + j] = old[src * 37 + j] - l * old[bi * 37 + j];
            }
        }
        __syncthreads();
        p ^= 1;
    }
    if (tid < 18) {
        const int c = tid;
        double x[18];
#pragma unroll
        for (int i = 17; i >= 0; --i) {
            double s = W->Y[i * 19 + c];
#pragma unroll
            for (int j = i + 1; j < 18; ++j) s = s - W->U[i * 19 + j] * x[j];
            x[i] = s / W->U[i * 19 + i];
        }
#pragma unroll
        for (int i = 0; i < 18; ++i) inv_out[i * 18 + c] = x[i];
    }
    __syncthreads();
}

__device__ int g_lu_variant = 3;   // 0: redundant (one barrier per step), 1: lean (three barriers, no redundant f64 work), 2: lean + rolled back substitution,
                                   // 3 (default): lean + REDUX pivot search -- parity-tested on a B200, pivot phase 1.5 k -> 0.35 k cycles per step
__device__ __forceinline__ void block_lu_inverse18(const double* a_in, InvScratch* W, double* inv_out) {
    const int v = g_lu_variant;
    if (v == 0) block_lu_inverse18_redundant(a_in, W, inv_out);
    else block_lu_inverse18_lean(a_in, W, inv_out, v == 2, v == 3);
}

// same arithmetic as ieskf_solve (lio_core.cuh), block-cooperative with the two LU inverses done by warp 0
__device__ __noinline__ void ieskf_solve_block(const LioParams& P, LioCtrl* ctrl, int iter, SolveScratch* S, InvScratch* W = nullptr) {
    const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: no divergence handling around the LU shuffles
    double* state = ctrl->state;
    double* cov = state + 24;
    for (int e = tid; e < 29; e += nthreads) {
        const long long hi = (long long)*(volatile unsigned long long*)&ctrl->acc[iter][2 * e];
        const long long lo = (long long)*(volatile unsigned long long*)&ctrl->acc[iter][2 * e + 1];
        const double v = fx_value(hi, lo);
        if (e < 21) {
            int ei = 0, base = 0;
            while (e >= base + (6 - ei)) { base += 6 - ei; ++ei; }
            const int ej = ei + (e - base);
            S->HTH[ei * 6 + ej] = v;
            S->HTH[ej * 6 + ei] = v;
        } else if (e < 27) {
            S->HTz[e - 21] = v;
        } else if (e == 27) {
            ctrl->stats[iter].total_residual = v;
        } else {
            ctrl->stats[iter].n_match = (double)((hi << 32) + lo);
        }
    }
    __syncthreads();
    // state.cov.inverse() is the same matrix in every iteration of a scan (the covariance only changes at the end):
    // it is computed once per scan by k_pinv and read here
    for (int idx = tid; idx < 324; idx += nthreads) {
        const int i = idx / 18, j = idx % 18;
        const double hth = (i < 6 && j < 6) ? S->HTH[i * 6 + j] : 0.0;
        S->a[idx] = hth + ctrl->Pinv[idx];
    }
    __syncthreads();
    if (W) {
        block_lu_inverse18(S->a, W, S->K1);          // whole block (kernels launched with INV_THREADS threads)
    } else {
        if (warp == 0) warp_lu_inverse18(S->a, S->lu, S->piv, S->K1, lane);   // fused-solve tail of a 128-thread residual block
        __syncthreads();
    }
    for (int idx = tid; idx < 18 * 6; idx += nthreads) {
        const int i = idx / 6, j = idx % 6;
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s = s + S->K1[i * 18 + k] * S->HTH[k * 6 + j];
        ctrl->G[i * 18 + j] = s;
    }
    if (tid == 0) state_minus(ctrl->state_prop, state, S->vec);
    __syncthreads();
    for (int i = tid; i < 18; i += nthreads) {
        double s1 = 0.0, s2 = 0.0;
        for (int k = 0; k < 6; ++k) s1 = s1 + S->K1[i * 18 + k] * S->HTz[k];
        for (int k = 0; k < 6; ++k) s2 = s2 + ctrl->G[i * 18 + k] * S->vec[k];
        S->sol[i] = (s1 + S->vec[i]) - s2;
    }
    __syncthreads();
    if (tid == 0) {
        state_plus(state, S->sol);
        const double* sol = S->sol;
        const double rn = sqrt((sol[0] * sol[0] + sol[1] * sol[1]) + sol[2] * sol[2]);
        const double tn = sqrt((sol[3] * sol[3] + sol[4] * sol[4]) + sol[5] * sol[5]);
        const int converged = ((rn * 57.3 < 0.01) && (tn * 100 < 0.015)) ? 1 : 0;
        IterStats& st = ctrl->stats[iter];
        for (int i = 0; i < 36; ++i) st.HTH[i] = S->HTH[i];
        for (int i = 0; i < 6; ++i) st.HTz[i] = S->HTz[i];
        for (int i = 0; i < 18; ++i) st.solution[i] = sol[i];
        st.converged = converged;
        ctrl->iters_run = iter + 1;
        int rematch = ctrl->rematch_num;
        if (converged || ((rematch == 0) && (iter == P.max_iter - 2))) rematch++;
        ctrl->rematch_num = rematch;
        S->flags[0] = (rematch >= 2 || iter == P.max_iter - 1) ? 1 : 0;
    }
    __syncthreads();
    if (S->flags[0]) {
        for (int idx = tid; idx < 324; idx += nthreads) {
            const int i = idx / 18, j = idx % 18;
            double s = 0.0;
            for (int k = 0; k < 18; ++k) {
                const double ig = ((i == k) ? 1.0 : 0.0) - ((k < 6) ? ctrl->G[i * 18 + k] : 0.0);
                s = s + ig * cov[k * 18 + j];
            }
            S->ncov[idx] = s;
        }
        __syncthreads();
        for (int idx = tid; idx < 324; idx += nthreads) cov[idx] = S->ncov[idx];
        if (tid == 0) ctrl->stop = 1;
    }
    __syncthreads();
}

// K2+K3: one thread per scan point; the block's 30 fixed-point sums are reduced with warp shuffles
// (integer adds: exact, order-free) and folded into the iteration's global accumulators with 60 atomics.
#define RES_THREADS 128
__global__ void __launch_bounds__(RES_THREADS) k_residual(VoxelMapDev map, LioParams P, ScanBuf sb, LioCtrl* ctrl, int iter, int n, int fused_solve) {
    __shared__ double s_state[24 + 6 * 18];
    __shared__ long long s_part[RES_THREADS / 32][IM_NTERMS];
    if (ctrl->stop) return;
    // stage rot/pos and the 6x6 pose block of the covariance (the only parts of the state this pass reads)
    for (int i = threadIdx.x; i < 24 + 6 * 18; i += blockDim.x) s_state[i] = ctrl->state[i];
    __syncthreads();
    long long acc[IM_NTERMS];
#pragma unroll
    for (int k = 0; k < IM_NTERMS; ++k) acc[k] = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        long long t[IM_NTERMS];
        if (residual_point(map, P, sb, s_state, i, t, map.err)) {
#pragma unroll
            for (int k = 0; k < IM_NTERMS; ++k) acc[k] += t[k];
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < IM_NTERMS - 1; ++k) {
        long long v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < IM_NTERMS - 1) {
        long long v = 0;
        for (int w = 0; w < RES_THREADS / 32; ++w) v += s_part[w][threadIdx.x];
        if (v != 0) {
            atomicAdd(&ctrl->acc[iter][2 * threadIdx.x], (unsigned long long)(v >> 32));
            atomicAdd(&ctrl->acc[iter][2 * threadIdx.x + 1], (unsigned long long)(v & 0xffffffffLL));
        }
    }
    if (!fused_solve) return;
    // the last block to publish its sums runs the 18x18 solve and the state update of this iteration
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ctrl->blocks_done[iter], 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {
        __shared__ SolveScratch S;
        __threadfence();
        ieskf_solve_block(P, ctrl, iter, &S);
    }
}

// sharded VoxelMap: pass 1 (match where this rank owns the voxel, publish bits) and pass 2 (terms + integer reduction)
__global__ void __launch_bounds__(RES_THREADS) k_shard_pass1(VoxelMapDev map, LioParams P, ScanBuf sb, LioCtrl* ctrl, int n, unsigned int* bits, int words) {
    __shared__ double s_state[24 + 6 * 18];
    if (ctrl->stop) return;
    for (int i = threadIdx.x; i < 24 + 6 * 18; i += blockDim.x) s_state[i] = ctrl->state[i];
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) shard_pass1_point(map, P, sb, s_state, i, bits, bits + words);
}
__global__ void __launch_bounds__(RES_THREADS) k_shard_pass2(VoxelMapDev map, LioParams P, ScanBuf sb, LioCtrl* ctrl, int iter, int n, const unsigned int* bits, int words) {
    __shared__ double s_state[24 + 6 * 18];
    __shared__ long long s_part[RES_THREADS / 32][IM_NTERMS];
    if (ctrl->stop) return;
    for (int i = threadIdx.x; i < 24 + 6 * 18; i += blockDim.x) s_state[i] = ctrl->state[i];
    __syncthreads();
    long long acc[IM_NTERMS];
#pragma unroll
    for (int k = 0; k < IM_NTERMS; ++k) acc[k] = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        long long t[IM_NTERMS];
        if (shard_pass2_point(map, P, sb, s_state, i, bits, bits + words, t, map.err)) {
#pragma unroll
            for (int k = 0; k < IM_NTERMS; ++k) acc[k] += t[k];
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < IM_NTERMS - 1; ++k) {
        long long v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < IM_NTERMS - 1) {
        long long v = 0;
        for (int w = 0; w < RES_THREADS / 32; ++w) v += s_part[w][threadIdx.x];
        if (v != 0) {
            atomicAdd(&ctrl->acc[iter][2 * threadIdx.x], (unsigned long long)(v >> 32));
            atomicAdd(&ctrl->acc[iter][2 * threadIdx.x + 1], (unsigned long long)(v & 0xffffffffLL));
        }
    }
}

// ---- sharded VoxelMap over peer windows (peer_win.cuh): the two exchanges of an iteration are fused into the kernels.
// Window of one rank:  flag[2][8] u64 | acc[8][IM_NTERMS*2] u64 (one row per source rank) | bits[8][2][words_cap] u32
struct LioPeers {
    unsigned char* w[IM_MAX_RANKS];
    int rank, n, words_cap, pad;
};
#define IM_LIOWIN_ACC_OFF 128
#define IM_LIOWIN_BITS_OFF 4096
__device__ __forceinline__ unsigned long long* liowin_flag(unsigned char* w, int phase, int src) { return (unsigned long long*)w + phase * IM_MAX_RANKS + src; }
__device__ __forceinline__ unsigned long long* liowin_acc(unsigned char* w, int src) { return (unsigned long long*)(w + IM_LIOWIN_ACC_OFF) + src * (IM_NTERMS * 2); }
__device__ __forceinline__ unsigned int* liowin_bits(unsigned char* w, int src, int which, int words_cap) { return (unsigned int*)(w + IM_LIOWIN_BITS_OFF) + (size_t)(src * 2 + which) * words_cap; }

// pass 1 + publish: every warp ballots the two bits of its 32 consecutive points and lane r stores the two words into rank
// r's window (row = this rank); the last block raises this rank's pass-1 flag in every peer window.
__global__ void __launch_bounds__(RES_THREADS) k_shard_pass1_p2p(VoxelMapDev map, LioParams P, ScanBuf sb, LioCtrl* ctrl, int n, LioPeers pe, unsigned long long epoch) {
    __shared__ double s_state[24 + 6 * 18];
    __shared__ int s_last;
    if (ctrl->stop) return;
    for (int i = threadIdx.x; i < 24 + 6 * 18; i += blockDim.x) s_state[i] = ctrl->state[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n32 = (n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += gridDim.x * blockDim.x) {
        bool ex = false, ok1 = false;
        if (i < n) shard_pass1_flags(map, P, sb, s_state, i, &ex, &ok1);
        const unsigned int we = __ballot_sync(0xffffffffu, ex), wo = __ballot_sync(0xffffffffu, ok1);
        if (lane < pe.n) {
            liowin_bits(pe.w[lane], pe.rank, 0, pe.words_cap)[i >> 5] = we;
            liowin_bits(pe.w[lane], pe.rank, 1, pe.words_cap)[i >> 5] = wo;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ctrl->shard_cnt[0], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        if (threadIdx.x == 0) ctrl->shard_cnt[0] = 0;
        __threadfence_system();
        if (threadIdx.x < pe.n && threadIdx.x != pe.rank) immesh::st_release_sys(liowin_flag(pe.w[threadIdx.x], 0, pe.rank), epoch);
    }
}
// wait for the peers' bits + pass 2 + integer block reduction + publish: the last block copies this rank's 60 partial sums
// into row `rank` of every window and raises the pass-2 flag.
__global__ void __launch_bounds__(RES_THREADS) k_shard_pass2_p2p(VoxelMapDev map, LioParams P, ScanBuf sb, LioCtrl* ctrl, int iter, int n, LioPeers pe, unsigned long long epoch) {
    __shared__ double s_state[24 + 6 * 18];
    __shared__ long long s_part[RES_THREADS / 32][IM_NTERMS];
    __shared__ int s_last;
    if (ctrl->stop) return;
    if (threadIdx.x < pe.n && threadIdx.x != pe.rank) {
        immesh::wait_epoch(liowin_flag(pe.w[pe.rank], 0, threadIdx.x), epoch, map.err, IM_ERR_PEER_TIMEOUT);
    }
    for (int i = threadIdx.x; i < 24 + 6 * 18; i += blockDim.x) s_state[i] = ctrl->state[i];
    __syncthreads();
    long long acc[IM_NTERMS];
#pragma unroll
    for (int k = 0; k < IM_NTERMS; ++k) acc[k] = 0;
    unsigned char* mine = pe.w[pe.rank];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned int we = 0, wo = 0;
        for (int r = 0; r < pe.n; ++r) {
            we |= __ldcg(liowin_bits(mine, r, 0, pe.words_cap) + (i >> 5));
            wo |= __ldcg(liowin_bits(mine, r, 1, pe.words_cap) + (i >> 5));
        }
        long long t[IM_NTERMS];
        if (shard_pass2_flags(map, P, sb, s_state, i, (we >> (i & 31)) & 1u, (wo >> (i & 31)) & 1u, t, map.err)) {
#pragma unroll
            for (int k = 0; k < IM_NTERMS; ++k) acc[k] += t[k];
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < IM_NTERMS - 1; ++k) {
        long long v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < IM_NTERMS - 1) {
        long long v = 0;
        for (int w = 0; w < RES_THREADS / 32; ++w) v += s_part[w][threadIdx.x];
        if (v != 0) {
            atomicAdd(&ctrl->acc[iter][2 * threadIdx.x], (unsigned long long)(v >> 32));
            atomicAdd(&ctrl->acc[iter][2 * threadIdx.x + 1], (unsigned long long)(v & 0xffffffffLL));
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ctrl->shard_cnt[1], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (threadIdx.x == 0) ctrl->shard_cnt[1] = 0;
        if (threadIdx.x < IM_NTERMS * 2) {
            const unsigned long long v = *(volatile unsigned long long*)&ctrl->acc[iter][threadIdx.x];
            for (int r = 0; r < pe.n; ++r) liowin_acc(pe.w[r], pe.rank)[threadIdx.x] = v;
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < pe.n && threadIdx.x != pe.rank) immesh::st_release_sys(liowin_flag(pe.w[threadIdx.x], 1, pe.rank), epoch);
    }
}
// wait for the peers' sums, add the rows (integers: order-free, identical on every rank), then the usual solve
__global__ void __launch_bounds__(INV_THREADS) k_solve_warp_p2p(LioParams P, LioCtrl* ctrl, int iter, LioPeers pe, unsigned long long epoch, int* err) {
    __shared__ SolveScratch S;
    __shared__ InvScratch W;
    if (ctrl->stop) return;
    if (threadIdx.x < pe.n && threadIdx.x != pe.rank) {
        immesh::wait_epoch(liowin_flag(pe.w[pe.rank], 1, threadIdx.x), epoch, err, IM_ERR_PEER_TIMEOUT);
    }
    __syncthreads();
    if (threadIdx.x < IM_NTERMS * 2) {
        unsigned long long v = 0;
        for (int r = 0; r < pe.n; ++r) v += __ldcg(liowin_acc(pe.w[pe.rank], r) + threadIdx.x);
        ctrl->acc[iter][threadIdx.x] = v;
    }
    __threadfence_block();
    __syncthreads();
    ieskf_solve_block(P, ctrl, iter, &S, &W);
}

#define SOLVE_THREADS 352
__global__ void __launch_bounds__(SOLVE_THREADS) k_solve(LioParams P, LioCtrl* ctrl, int iter) {
    __shared__ SolveScratch S;
    ieskf_solve(P, ctrl, iter, &S, threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(INV_THREADS) k_pinv(LioCtrl* ctrl) {
    __shared__ double a[324], inv[324];
    __shared__ InvScratch W;
    const long long t_in = clock64();
    for (int i = threadIdx.x; i < 324; i += blockDim.x) a[i] = ctrl->state[24 + i];
    __syncthreads();
    block_lu_inverse18(a, &W, inv);
    for (int i = threadIdx.x; i < 324; i += blockDim.x) ctrl->Pinv[i] = inv[i];
    if (threadIdx.x == 0) { g_inv_stamps[5] = t_in; g_inv_stamps[6] = clock64(); }
}
__global__ void __launch_bounds__(INV_THREADS) k_solve_warp(LioParams P, LioCtrl* ctrl, int iter) {
    __shared__ SolveScratch S;
    __shared__ InvScratch W;
    if (ctrl->stop) return;
    ieskf_solve_block(P, ctrl, iter, &S, &W);
}

__global__ void __launch_bounds__(SOLVE_THREADS) k_predict(LioCtrl* ctrl, double dt, double cov_gyr, double cov_acc) {
    __shared__ double T[324], Fx[324];
    predict_const_vel(ctrl->state, dt, cov_gyr, cov_acc, T, Fx, threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(128) k_grow_point(VoxelMapDev map, LioParams P, ScanBuf sb, LioCtrl* ctrl, int n, int mode) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) grow_point(map, P, sb, ctrl->state, i, mode);
}
__global__ void __launch_bounds__(128) k_grow_segment(ScanBuf sb) {
    const int nt = *sb.n_touched;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) grow_segment(sb, t);
}
__global__ void __launch_bounds__(128) k_grow_scatter(ScanBuf sb, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) grow_scatter(sb, i);
}
// one warp per touched root voxel, voxels claimed dynamically (their cost varies by orders of magnitude)
__global__ void __launch_bounds__(128) k_grow_voxel(VoxelMapDev map, LioParams P, ScanBuf sb, int mode, int* sorted_scratch, int* work_counter) {
    const int lane = threadIdx.x & 31;
    const int nt = *sb.n_touched;
    while (true) {
        int t = 0;
        if (lane == 0) t = atomicAdd(work_counter, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= nt) break;
        grow_voxel(map, P, sb, t, mode, lane, 32, sorted_scratch);
    }
}
__global__ void k_grow_finish(VoxelMapDev map, ScanBuf sb, int* work_counter) {
    recycle_chunks(map, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) {
        *sb.n_touched = 0;
        *sb.seg_top = 0;
        *work_counter = 0;
    }
}
// BuildResidualListOMP on caller-supplied Point_with_var data (world point + covariance per point)
__global__ void __launch_bounds__(128) k_match_pv(VoxelMapDev map, LioParams P, ScanBuf sb, const double* pw, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const MatchResult mr = match_point(map, P, pw + (size_t)i * 3, sb.var + (size_t)i * 6);
        sb.match_node[i] = mr.node;
        sb.match_layer[i] = mr.layer;
    }
}
// ptpl payload of every matched point (diagnostic / drop-in immesh_residual_build)
__global__ void k_gather_ptpl(VoxelMapDev map, ScanBuf sb, int n, double* out /*[n][31]*/) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int nd = sb.match_node[i];
        if (nd < 0) continue;
        const PlaneRec& pl = map.planes[nd];
        double* o = out + (size_t)i * 31;
        for (int j = 0; j < 3; ++j) { o[j] = (double)sb.body[i * 3 + j]; o[3 + j] = pl.normal[j]; o[6 + j] = pl.center[j]; }
        o[9] = (double)pl.d;
        for (int j = 0; j < 21; ++j) o[10 + j] = pl.pv[j];
    }
}

// ------------------------------------------------------------------ host side

#include "nccl_api.hpp"
using immesh::nccl; using immesh::nccl_uid_t; using immesh::nccl_comm_t; using immesh::kNcclUint32; using immesh::kNcclUint64; using immesh::kNcclSum;

static void fill_params(const immesh_lio_config* c, LioParams& P) {
    P.voxel_size = c->voxel_size;
    P.voxel_size_f = (float)c->voxel_size;
    P.voxel_size_ins = (double)P.voxel_size_f;
    P.max_layer = c->max_layer;
    for (int i = 0; i < 5; ++i) P.layer_init[i] = c->layer_init_size[i];
    P.max_points = c->max_points_size;
    P.planer_threshold = (float)c->min_eigen_value;
    P.dept_err = (float)c->dept_err;
    // DEG2RAD is PCL's macro ((x)*0.017453293); the squared sine is a per-configuration constant
    const float be = (float)c->beam_err;
    const double s = std::sin((double)be * 0.017453293);
    P.dir_var = s * s;
    const double sc = std::sin((double)(float)0.01 * 0.017453293);  // CALIB_ANGLE_COV, include/common_lib.h:41
    P.dir_var_calib = sc * sc;
    P.calib_laser = c->calib_laser;
    P.max_iter = c->max_iteration;
    for (int i = 0; i < 9; ++i) P.extR[i] = c->ext_R[i];
    for (int i = 0; i < 3; ++i) P.extT[i] = c->ext_T[i];
}

template <class T>
static cudaError_t dev_alloc(immesh_lio* h, T** p, size_t count, int memset_byte = -1) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) return e;
    h->allocs.push_back(q);
    *p = (T*)q;
    if (memset_byte >= 0) e = cudaMemset(q, memset_byte, count * sizeof(T));
    return e;
}

static int grid_for(const immesh_lio* h, int n, int threads, int max_waves = 8) {
    int g = (n + threads - 1) / threads;
    const int cap = h->n_sm * max_waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return g;
}

extern "C" {

int immesh_lio_create(const immesh_lio_config* cfg, immesh_lio_t** out) {
    if (!cfg || !out) return im_fail(IMMESH_E_INVALID, "null argument");
    if (cfg->max_iteration < 1 || cfg->max_iteration > IM_MAX_ITER) return im_fail(IMMESH_E_INVALID, "max_iteration must be in [1,8]");
    if (cfg->max_layer < 0 || cfg->max_layer > 4) return im_fail(IMMESH_E_INVALID, "max_layer must be in [0,4]");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return im_fail(IMMESH_E_NO_DEVICE, "no CUDA device: immesh_b200 has no CPU path");
    immesh_lio* h = new immesh_lio();
    fill_params(cfg, h->P);
    h->bps = std::getenv("IMMESH_LIO_BPS") ? std::atoi(std::getenv("IMMESH_LIO_BPS")) : 4;
    h->use_graph = std::getenv("IMMESH_GRAPH") ? std::atoi(std::getenv("IMMESH_GRAPH")) : 1;
    h->fused_solve = std::getenv("IMMESH_FUSED_SOLVE") ? std::atoi(std::getenv("IMMESH_FUSED_SOLVE")) : 0;
    const int caplog = cfg->hash_capacity_log2 ? cfg->hash_capacity_log2 : 22;
    h->cap = (size_t)1 << caplog;
    h->max_nodes = cfg->max_nodes ? cfg->max_nodes : (4 << 20);
    h->max_chunks = cfg->max_chunks ? cfg->max_chunks : (4 << 20);
    h->max_scan = cfg->max_scan_points ? cfg->max_scan_points : (2 << 20);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, dev);
    IM_CUDA(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, im_stream_priority("IMMESH_LIO_PRIO")));
    for (auto& e : h->ev) IM_CUDA(cudaEventCreate(&e));
    IM_CUDA(cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, im_stream_priority("IMMESH_LIO_PRIO")));
    IM_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    IM_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) IM_CUDA(cudaEventCreateWithFlags(&h->ev_slot[i], cudaEventDisableTiming));
    VoxelMapDev& m = h->map;
    IM_CUDA(dev_alloc(h, &m.keys, h->cap, 0xFF));
    IM_CUDA(dev_alloc(h, &m.root_node, h->cap, 0xFF));
    m.cap_mask = (unsigned)(h->cap - 1);
    IM_CUDA(dev_alloc(h, &m.nodes, (size_t)h->max_nodes));
    IM_CUDA(dev_alloc(h, &m.planes, (size_t)h->max_nodes));
    IM_CUDA(dev_alloc(h, &m.chunks, (size_t)h->max_chunks));
    IM_CUDA(dev_alloc(h, &m.avail, (size_t)h->max_chunks));
    IM_CUDA(dev_alloc(h, &m.pending, (size_t)h->max_chunks));
    IM_CUDA(dev_alloc(h, &h->d_counters, 16, 0));
    m.node_count = h->d_counters + 0; m.chunk_bump = h->d_counters + 1; m.avail_top = h->d_counters + 2; m.pending_n = h->d_counters + 3;
    m.err = h->d_counters + 4; m.n_roots = h->d_counters + 5;
    m.max_nodes = h->max_nodes; m.max_chunks = h->max_chunks;
    ScanBuf& sb = h->sb;
    const size_t ms = (size_t)h->max_scan;
    float* d_body = nullptr;
    IM_CUDA(dev_alloc(h, &d_body, ms * 3));
    sb.body = d_body;
    h->d_body_own = d_body;
    IM_CUDA(dev_alloc(h, &sb.body_cov, ms * 6));
    IM_CUDA(dev_alloc(h, &sb.p_imu, ms * 3));
    IM_CUDA(dev_alloc(h, &sb.bv_imu, ms * 6));
    IM_CUDA(dev_alloc(h, &sb.match_node, ms, 0xFF));
    IM_CUDA(dev_alloc(h, &sb.match_layer, ms, 0));
    IM_CUDA(dev_alloc(h, &sb.pw, ms * 3));
    IM_CUDA(dev_alloc(h, &sb.var, ms * 6));
    IM_CUDA(dev_alloc(h, &sb.sortkey, ms));
    IM_CUDA(dev_alloc(h, &sb.slot, ms));
    IM_CUDA(dev_alloc(h, &sb.seg, ms));
    IM_CUDA(dev_alloc(h, &h->d_sorted, ms));
    IM_CUDA(dev_alloc(h, &sb.touched, ms));
    IM_CUDA(dev_alloc(h, &sb.slot_count, h->cap, 0));
    IM_CUDA(dev_alloc(h, &sb.slot_offset, h->cap, 0));
    IM_CUDA(dev_alloc(h, &sb.slot_cursor, h->cap, 0));
    sb.n_touched = h->d_counters + 6; sb.seg_top = h->d_counters + 7;
    sb.n = 0;
    IM_CUDA(dev_alloc(h, &h->d_ctrl, 1, 0));
    IM_CUDA(cudaMallocHost((void**)&h->h_body, 2 * ms * 3 * sizeof(float)));
    IM_CUDA(cudaMallocHost((void**)&h->h_state, 2 * (IM_STATE_DOUBLES + 64) * sizeof(double)));
    IM_CUDA(cudaMallocHost((void**)&h->h_ints, 64 * sizeof(int)));
    // StatesGroup(): identity rotation, cov = INIT_COV * I  (include/common_lib.h:201-211)
    std::memset(h->h_state, 0, IM_STATE_DOUBLES * sizeof(double));
    h->h_state[0] = h->h_state[4] = h->h_state[8] = 1.0;
    for (int i = 0; i < 18; ++i) h->h_state[24 + i * 18 + i] = 0.0000001;
    IM_CUDA(cudaMemcpy(h->d_ctrl->state, h->h_state, IM_STATE_DOUBLES * sizeof(double), cudaMemcpyHostToDevice));
    IM_CUDA(cudaDeviceSynchronize());
    if (const char* v = std::getenv("IMMESH_LU_VARIANT")) {   // experiments: 0 = redundant-pivot variant, 1 = lean (shuffle tournament), 2 = rolled back substitution, 3 = default
        const int iv = std::atoi(v);
        IM_CUDA(cudaMemcpyToSymbol(g_lu_variant, &iv, sizeof(int)));
    }
    *out = h;
    return IMMESH_OK;
}

int immesh_lio_destroy(immesh_lio_t* h) {
    if (!h) return IMMESH_OK;
    cudaStreamSynchronize(h->stream);
    h->graph.destroy();
    if (h->win.local) immesh::peer_window_close(h->win);
    if (h->nccl_comm && nccl().CommDestroy) nccl().CommDestroy(h->nccl_comm);
    for (void* p : h->allocs) cudaFree(p);
    if (h->h_body) cudaFreeHost(h->h_body);
    if (h->h_state) cudaFreeHost(h->h_state);
    if (h->h_ints) cudaFreeHost(h->h_ints);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (int i = 0; i < 2; ++i) if (h->ev_slot[i]) cudaEventDestroy(h->ev_slot[i]);
    delete h;
    return IMMESH_OK;
}

// ---- multi-GPU: shard the VoxelMap of this handle over `nranks` processes (one GPU each, NCCL over NVLink)
int immesh_comm_unique_id(char* out128) {
    if (!out128) return im_fail(IMMESH_E_INVALID, "null argument");
    if (!nccl().load()) return im_fail(IMMESH_E_CUDA, "libnccl.so.2 not found");
    nccl_uid_t id;
    if (nccl().GetUniqueId(&id)) return im_fail(IMMESH_E_CUDA, "ncclGetUniqueId failed");
    std::memcpy(out128, id.internal, 128);
    return IMMESH_OK;
}
int immesh_lio_shard(immesh_lio_t* h, int rank, int nranks, const char* unique_id128) {
    if (!h || !unique_id128 || nranks < 1 || rank < 0 || rank >= nranks) return im_fail(IMMESH_E_INVALID, "bad argument");
    if (nranks == 1) { h->P.shard_rank = 0; h->P.shard_n = 1; return IMMESH_OK; }
    if (!nccl().load()) return im_fail(IMMESH_E_CUDA, "libnccl.so.2 not found");
    nccl_uid_t id;
    std::memcpy(id.internal, unique_id128, 128);
    nccl_comm_t comm = nullptr;
    const int rc = nccl().CommInitRank(&comm, nranks, id, rank);
    if (rc) return im_fail(IMMESH_E_CUDA, nccl().GetErrorString ? nccl().GetErrorString(rc) : "ncclCommInitRank failed");
    h->nccl_comm = comm;
    h->P.shard_rank = rank;
    h->P.shard_n = nranks;
    if (!h->d_bits) IM_CUDA(dev_alloc(h, &h->d_bits, (size_t)2 * (h->max_scan / 32 + 2), 0));
    // peer window over NVLink (CUDA IPC): the exchanges of the residual pass are then fused into its kernels.  IMMESH_SHARD_NCCL=1
    // keeps the NCCL all-reduce sequence (the baseline the fused path is measured against).
    const char* force_nccl = std::getenv("IMMESH_SHARD_NCCL");
    if (!(force_nccl && force_nccl[0] == '1')) {
        h->words_cap = h->max_scan / 32 + 2;
        const size_t bytes = IM_LIOWIN_BITS_OFF + (size_t)IM_MAX_RANKS * 2 * h->words_cap * sizeof(unsigned int);
        const cudaError_t e = immesh::peer_window_open(h->win, bytes, rank, nranks, comm, h->stream);
        if (e != cudaSuccess) {
            std::fprintf(stderr, "[immesh_b200] rank %d: peer window unavailable (%s); sharded residual pass falls back to NCCL all-reduces\n", rank, cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    return IMMESH_OK;
}
// diagnostic: clock64 stamps of the last 18x18 inverse (k_pinv): [enter, after init, after step 0, after LU, after back substitution, kernel start, kernel end]
int immesh_debug_inverse_stamps(long long* out8 /*[16]*/) {
    if (!out8) return im_fail(IMMESH_E_INVALID, "null argument");
    IM_CUDA(cudaDeviceSynchronize());
    IM_CUDA(cudaMemcpyFromSymbol(out8, g_inv_stamps, 16 * sizeof(long long)));
    return IMMESH_OK;
}
int immesh_lio_shard_transport(immesh_lio_t* h) {   // 0 = not sharded, 1 = NCCL all-reduces, 2 = fused peer-window exchange
    if (!h) return 0;
    if (h->P.shard_n <= 1) return 0;
    return h->win.ok ? 2 : 1;
}

int immesh_lio_set_state(immesh_lio_t* h, const double* s) {
    if (!h || !s) return im_fail(IMMESH_E_INVALID, "null argument");
    std::memcpy(h->h_state, s, IM_STATE_DOUBLES * sizeof(double));
    IM_CUDA(cudaMemcpyAsync(h->d_ctrl->state, h->h_state, IM_STATE_DOUBLES * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    IM_CUDA(cudaStreamSynchronize(h->stream));
    return IMMESH_OK;
}
int immesh_lio_get_state(immesh_lio_t* h, double* s) {
    if (!h || !s) return im_fail(IMMESH_E_INVALID, "null argument");
    IM_CUDA(cudaMemcpyAsync(h->h_state, h->d_ctrl->state, IM_STATE_DOUBLES * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaStreamSynchronize(h->stream));
    std::memcpy(s, h->h_state, IM_STATE_DOUBLES * sizeof(double));
    return IMMESH_OK;
}

static int upload_scan(immesh_lio* h, const float* body, int n, int on_device = 0, int slot = 0) {
    if ((!body && n > 0) || n < 0) return im_fail(IMMESH_E_INVALID, "bad scan");
    if (n > h->max_scan) return im_fail(IMMESH_E_CAPACITY, "scan larger than max_scan_points");
    if (on_device) {
        h->sb.body = body;  // caller-owned device buffer, must stay valid until the next call on this handle
    } else {
        h->sb.body = h->d_body_own;
        if (n > 0) {
            float* stage = h->h_body + (size_t)slot * h->max_scan * 3;
            std::memcpy(stage, body, (size_t)n * 3 * sizeof(float));
            IM_CUDA(cudaMemcpyAsync((void*)h->d_body_own, stage, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        }
    }
    h->sb.n = n;
    h->last_n = n;
    return IMMESH_OK;
}
static int check_flags(immesh_lio* h) {
    IM_CUDA(cudaMemcpyAsync(h->h_ints, h->d_counters, 16 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaStreamSynchronize(h->stream));
    const int err = h->h_ints[4];
    if (err & (IM_ERR_NODE_POOL | IM_ERR_CHUNK_POOL | IM_ERR_HASH_FULL | IM_ERR_SEG_POOL)) return im_fail(IMMESH_E_CAPACITY, "device pool overflow (raise the capacities in immesh_lio_config)");
    if (err & (IM_ERR_KEY_RANGE | IM_ERR_FX_RANGE)) return im_fail(IMMESH_E_RANGE, "coordinate / normal-equation term outside the representable range");
    if (err & IM_ERR_PEER_TIMEOUT) return im_fail(IMMESH_E_CUDA, "sharded mode: a peer rank did not publish its data in time (peer window epoch flag)");
    return IMMESH_OK;
}
static void launch_grow(immesh_lio* h, int n, int mode) {
    if (n <= 0) return;
    IM_LAUNCH(k_grow_point, grid_for(h, n, 128), 128, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, n, mode);
    IM_LAUNCH(k_grow_segment, grid_for(h, n, 128, 2), 128, 0, h->stream, h->sb);
    IM_LAUNCH(k_grow_scatter, grid_for(h, n, 128), 128, 0, h->stream, h->sb, n);
    IM_LAUNCH(k_grow_voxel, h->n_sm * h->bps, 128, 0, h->stream, h->map, h->P, h->sb, mode, h->d_sorted, h->d_counters + 8);
    IM_LAUNCH(k_grow_finish, 1, 256, 0, h->stream, h->map, h->sb, h->d_counters + 8);
}
// sharded VoxelMap: per iteration  pass1 -> all-reduce(bit words) -> pass2 -> all-reduce(58 int64 sums) -> solve (every rank)
static int launch_estimate_sharded(immesh_lio* h, int n) {
    const int words = (n + 31) / 32 + 1;
    cudaEventRecord(h->ev_fork, h->stream);
    cudaStreamWaitEvent(h->stream2, h->ev_fork, 0);
    IM_LAUNCH(k_pinv, 1, INV_THREADS, 0, h->stream2, h->d_ctrl);
    cudaEventRecord(h->ev_join, h->stream2);
    IM_LAUNCH(k_reset_scan, 2, 256, 0, h->stream, h->sb, h->d_ctrl, 1);
    if (n > 0) IM_LAUNCH(k_prepare, grid_for(h, n, 128), 128, 0, h->stream, h->P, h->sb, n);
    const int g = grid_for(h, n > 0 ? n : 1, RES_THREADS, 4);
    for (int it = 0; it < h->P.max_iter; ++it) {
        IM_CUDA(cudaMemsetAsync(h->d_bits, 0, (size_t)2 * words * sizeof(unsigned int), h->stream));
        if (n > 0) IM_LAUNCH(k_shard_pass1, g, RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, n, h->d_bits, words);
        if (nccl().AllReduce(h->d_bits, h->d_bits, (size_t)2 * words, kNcclUint32, kNcclSum, h->nccl_comm, h->stream)) return im_fail(IMMESH_E_CUDA, "ncclAllReduce(bits) failed");
        if (n > 0) IM_LAUNCH(k_shard_pass2, g, RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, it, n, h->d_bits, words);
        if (nccl().AllReduce(&h->d_ctrl->acc[it][0], &h->d_ctrl->acc[it][0], (size_t)IM_NTERMS * 2, kNcclUint64, kNcclSum, h->nccl_comm, h->stream)) return im_fail(IMMESH_E_CUDA, "ncclAllReduce(acc) failed");
        if (it == 0) cudaStreamWaitEvent(h->stream, h->ev_join, 0);
        IM_LAUNCH(k_solve_warp, 1, INV_THREADS, 0, h->stream, h->P, h->d_ctrl, it);
    }
    return IMMESH_OK;
}
// sharded VoxelMap over peer windows: per iteration  pass1(+publish bits) -> pass2(wait bits, +publish sums) -> solve(wait sums)
static void launch_estimate_p2p(immesh_lio* h, int n) {
    const bool replay = immesh::im_replaying();
    if (n > 0) {
        if (!replay) { cudaEventRecord(h->ev_fork, h->stream); cudaStreamWaitEvent(h->stream2, h->ev_fork, 0); }
        IM_LAUNCH(k_pinv, 1, INV_THREADS, 0, h->stream2, h->d_ctrl);
        if (!replay) cudaEventRecord(h->ev_join, h->stream2);
    }
    IM_LAUNCH(k_reset_scan, 2, 256, 0, h->stream, h->sb, h->d_ctrl, 1);
    if (n <= 0) return;   // every rank sees the same (replicated) scan, so all of them skip the exchanges together
    IM_LAUNCH(k_prepare, grid_for(h, n, 128), 128, 0, h->stream, h->P, h->sb, n);
    LioPeers pe;
    for (int r = 0; r < IM_MAX_RANKS; ++r) pe.w[r] = h->win.peer[r];
    pe.rank = h->win.rank; pe.n = h->win.n; pe.words_cap = h->words_cap; pe.pad = 0;
    const int g = grid_for(h, n, RES_THREADS, 4);
    for (int it = 0; it < h->P.max_iter; ++it) {
        const unsigned long long e = h->win.epoch + 1 + it;   // one epoch per iteration (base advanced once per scan by the caller); the two phases have separate flag rows
        IM_LAUNCH(k_shard_pass1_p2p, g, RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, n, pe, e);
        IM_LAUNCH(k_shard_pass2_p2p, g, RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, it, n, pe, e);
        if (it == 0 && !replay) cudaStreamWaitEvent(h->stream, h->ev_join, 0);
        IM_LAUNCH(k_solve_warp_p2p, 1, INV_THREADS, 0, h->stream, h->P, h->d_ctrl, it, pe, e, h->map.err);
    }
}
static void launch_estimate(immesh_lio* h, int n) {
    if (h->P.shard_n > 1 && h->win.ok) { launch_estimate_p2p(h, n); return; }
    if (h->P.shard_n > 1) { launch_estimate_sharded(h, n); return; }
    const bool replay = immesh::im_replaying();   // graph replay: the fork/join edges are already part of the graph
    if (n > 0) {  // P^-1 on the side stream, overlapped with the scan preparation and the first residual pass
        if (!replay) { cudaEventRecord(h->ev_fork, h->stream); cudaStreamWaitEvent(h->stream2, h->ev_fork, 0); }
        IM_LAUNCH(k_pinv, 1, INV_THREADS, 0, h->stream2, h->d_ctrl);
        if (!replay) cudaEventRecord(h->ev_join, h->stream2);
    }
    IM_LAUNCH(k_reset_scan, 2, 256, 0, h->stream, h->sb, h->d_ctrl, 1);
    if (n <= 0) return;
    IM_LAUNCH(k_prepare, grid_for(h, n, 128), 128, 0, h->stream, h->P, h->sb, n);
    const int g = grid_for(h, n, RES_THREADS, 4);
    for (int it = 0; it < h->P.max_iter; ++it) {
        if (it == 0 && h->fused_solve && !replay) cudaStreamWaitEvent(h->stream, h->ev_join, 0);
        if (h->fused_solve) {
            IM_LAUNCH(k_residual, g, RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, it, n, 1);
        } else {
            IM_LAUNCH(k_residual, g, RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, it, n, 0);
            if (it == 0 && !replay) cudaStreamWaitEvent(h->stream, h->ev_join, 0);
            IM_LAUNCH(k_solve_warp, 1, INV_THREADS, 0, h->stream, h->P, h->d_ctrl, it);
        }
    }
}

int immesh_lio_predict(immesh_lio_t* h, double dt, double cov_gyr, double cov_acc) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    IM_LAUNCH(k_predict, 1, SOLVE_THREADS, 0, h->stream, h->d_ctrl, dt, cov_gyr, cov_acc);
    IM_CUDA(cudaGetLastError());
    IM_CUDA(cudaStreamSynchronize(h->stream));
    return IMMESH_OK;
}

int immesh_voxelmap_build(immesh_lio_t* h, const float* body, int n) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    int rc = upload_scan(h, body, n);
    if (rc) return rc;
    launch_grow(h, n, 1);
    IM_CUDA(cudaGetLastError());
    return check_flags(h);
}

int immesh_lio_estimate(immesh_lio_t* h, const float* body, int n, int* iters_run) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    int rc = upload_scan(h, body, n);
    if (rc) return rc;
    h->win.epoch += IM_MAX_ITER;   // epoch base of this scan (same sequence of calls on every rank)
    launch_estimate(h, n);
    IM_CUDA(cudaGetLastError());
    IM_CUDA(cudaMemcpyAsync(h->h_ints + 32, &h->d_ctrl->iters_run, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    rc = check_flags(h);
    if (iters_run) *iters_run = h->h_ints[32];
    return rc;
}

int immesh_voxelmap_update(immesh_lio_t* h) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    launch_grow(h, h->last_n, 0);
    IM_CUDA(cudaGetLastError());
    return check_flags(h);
}

static int lio_flags_status(int err) {
    if (err & (IM_ERR_NODE_POOL | IM_ERR_CHUNK_POOL | IM_ERR_HASH_FULL | IM_ERR_SEG_POOL)) return im_fail(IMMESH_E_CAPACITY, "device pool overflow (raise the capacities in immesh_lio_config)");
    if (err & (IM_ERR_KEY_RANGE | IM_ERR_FX_RANGE)) return im_fail(IMMESH_E_RANGE, "coordinate / normal-equation term outside the representable range");
    if (err & IM_ERR_PEER_TIMEOUT) return im_fail(IMMESH_E_CUDA, "sharded mode: a peer rank did not publish its data in time (peer window epoch flag)");
    return IMMESH_OK;
}
// queue predict + estimate + update for one scan; no host synchronisation unless both staging slots are busy
static int lio_enqueue(immesh_lio_t* h, const float* body, int n, int on_device, double dt, double cov_gyr, double cov_acc, bool allow_graph) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    const int s = (++h->step_counter) & 1;
    if (h->slot_busy[s]) { IM_CUDA(cudaEventSynchronize(h->ev_slot[s])); h->slot_busy[s] = 0; }
    int rc = upload_scan(h, body, n, on_device, s);
    if (rc) return rc;
    h->win.epoch += IM_MAX_ITER;   // epoch base of this scan: advanced exactly once per scan, also when the graph body runs twice
    // The pipelined entry points replay the scan's launch sequence as one CUDA graph (the sequence is host-launch-bound
    // otherwise); the blocking ones launch directly, with stage timing events in between.
    bool queued = false;
    if (allow_graph && h->use_graph && !profiler().enabled && (h->P.shard_n <= 1 || h->win.ok) && n > 0) {
        const unsigned sig = 1u | (dt > 0 ? 2u : 0u) | (h->fused_solve ? 4u : 0u) | (h->win.ok ? 8u : 0u) | ((unsigned)h->P.max_iter << 8);
        queued = immesh::run_graphed(h->graph, sig, h->stream, [&] {
            if (dt > 0) IM_LAUNCH(k_predict, 1, SOLVE_THREADS, 0, h->stream, h->d_ctrl, dt, cov_gyr, cov_acc);
            launch_estimate(h, n);
            launch_grow(h, n, 0);
        }) == cudaSuccess;
        if (!queued) h->use_graph = 0;   // not expected; keep working through direct launches
    }
    if (!queued) {
        IM_CUDA(cudaEventRecord(h->ev[0], h->stream));
        if (dt > 0) IM_LAUNCH(k_predict, 1, SOLVE_THREADS, 0, h->stream, h->d_ctrl, dt, cov_gyr, cov_acc);
        IM_CUDA(cudaEventRecord(h->ev[1], h->stream));
        launch_estimate(h, n);
        IM_CUDA(cudaEventRecord(h->ev[2], h->stream));
        launch_grow(h, n, 0);
    }
    IM_CUDA(cudaGetLastError());
    double* hs = h->h_state + (size_t)s * (IM_STATE_DOUBLES + 64);
    IM_CUDA(cudaMemcpyAsync(hs, h->d_ctrl->state, IM_STATE_DOUBLES * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaMemcpyAsync(h->h_ints + 32 + s, &h->d_ctrl->iters_run, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaMemcpyAsync(h->h_ints + 16 * s, h->d_counters, 16 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaEventRecord(h->ev[3], h->stream));
    IM_CUDA(cudaEventRecord(h->ev_slot[s], h->stream));
    h->slot_busy[s] = 1;
    return IMMESH_OK;
}
static int lio_wait_impl(immesh_lio_t* h, double* state_out, int* iters_run, bool timings) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    const int s = h->step_counter & 1;
    IM_CUDA(cudaStreamSynchronize(h->stream));
    h->slot_busy[0] = h->slot_busy[1] = 0;
    if (profiler().enabled) { cudaStreamSynchronize(h->stream2); profiler().collect(); }
    const double* hs = h->h_state + (size_t)s * (IM_STATE_DOUBLES + 64);
    if (state_out) std::memcpy(state_out, hs, IM_STATE_DOUBLES * sizeof(double));
    if (iters_run) *iters_run = h->h_ints[32 + s];
    if (timings) {
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, h->ev[0], h->ev[3]);
        cudaEventElapsedTime(&b, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&c, h->ev[2], h->ev[3]);
        h->last_ms[0] = a; h->last_ms[1] = b; h->last_ms[2] = c;
    }
    return lio_flags_status(h->h_ints[16 * s + 4]);
}

int immesh_lio_step(immesh_lio_t* h, const float* body, int n, double dt, double cov_gyr, double cov_acc, double* state_out, int* iters_run) {
    int rc = lio_enqueue(h, body, n, 0, dt, cov_gyr, cov_acc, false);
    return rc ? rc : lio_wait_impl(h, state_out, iters_run, true);
}
int immesh_lio_step_dev(immesh_lio_t* h, const float* d_body, int n, double dt, double cov_gyr, double cov_acc, double* state_out, int* iters_run) {
    int rc = lio_enqueue(h, d_body, n, 1, dt, cov_gyr, cov_acc, false);
    return rc ? rc : lio_wait_impl(h, state_out, iters_run, true);
}
int immesh_lio_step_async(immesh_lio_t* h, const float* body, int n, int on_device, double dt, double cov_gyr, double cov_acc) {
    return lio_enqueue(h, body, n, on_device, dt, cov_gyr, cov_acc, true);
}
int immesh_lio_wait(immesh_lio_t* h, double* state_out, int* iters_run) { return lio_wait_impl(h, state_out, iters_run, false); }
// queue a write of `bytes` bytes over a caller-provided device buffer on the localization stream (benchmark L2 flush)
int immesh_lio_enqueue_memset(immesh_lio_t* h, void* d_buf, size_t bytes) {
    if (!h || !d_buf) return im_fail(IMMESH_E_INVALID, "null argument");
    IM_CUDA(cudaMemsetAsync(d_buf, 0, bytes, h->stream));
    return IMMESH_OK;
}

int immesh_lio_last_timing(immesh_lio_t* h, double* ms) {
    if (!h || !ms) return im_fail(IMMESH_E_INVALID, "null argument");
    ms[0] = h->last_ms[0]; ms[1] = h->last_ms[1]; ms[2] = h->last_ms[2];
    return IMMESH_OK;
}

int immesh_residual_build(immesh_lio_t* h, const float* body, int n, int* index_layer, double* ptpl, int cap, int* n_out) {
    if (!h || !n_out) return im_fail(IMMESH_E_INVALID, "null argument");
    int rc = upload_scan(h, body, n);
    if (rc) return rc;
    if (!h->d_ptpl) IM_CUDA(dev_alloc(h, &h->d_ptpl, (size_t)h->max_scan * 31));
    IM_LAUNCH(k_reset_scan, 2, 256, 0, h->stream, h->sb, h->d_ctrl, 0);
    *n_out = 0;
    if (n == 0) return IMMESH_OK;
    IM_LAUNCH(k_prepare, grid_for(h, n, 128), 128, 0, h->stream, h->P, h->sb, n);
    IM_LAUNCH(k_residual, grid_for(h, n, RES_THREADS, 4), RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, 0, n, 0);
    IM_LAUNCH(k_gather_ptpl, grid_for(h, n, 128), 128, 0, h->stream, h->map, h->sb, n, h->d_ptpl);
    IM_CUDA(cudaGetLastError());
    std::vector<int> node(n), layer(n);
    std::vector<double> vals((size_t)n * 31);
    IM_CUDA(cudaMemcpyAsync(node.data(), h->sb.match_node, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaMemcpyAsync(layer.data(), h->sb.match_layer, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaMemcpyAsync(vals.data(), h->d_ptpl, (size_t)n * 31 * 8, cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaStreamSynchronize(h->stream));
    int m = 0;
    for (int i = 0; i < n; ++i) {
        if (node[i] < 0) continue;
        if (m < cap) {
            if (index_layer) { index_layer[2 * m] = i; index_layer[2 * m + 1] = layer[i]; }
            if (ptpl) std::memcpy(ptpl + (size_t)m * 31, vals.data() + (size_t)i * 31, 31 * 8);
        }
        ++m;
    }
    *n_out = m;
    return check_flags(h);
}

static void var9_to_6(const double* v9, double* v6) {  // upper triangle of the caller's (symmetric) 3x3
    v6[0] = v9[0]; v6[1] = v9[1]; v6[2] = v9[2]; v6[3] = v9[4]; v6[4] = v9[5]; v6[5] = v9[8];
}
static int upload_pv(immesh_lio* h, const double* pts_world, const double* var9, int n) {
    if ((!pts_world || !var9) && n > 0) return im_fail(IMMESH_E_INVALID, "null argument");
    if (n < 0 || n > h->max_scan) return im_fail(IMMESH_E_CAPACITY, "too many points");
    std::vector<float> pw((size_t)n * 3);
    std::vector<double> v6((size_t)n * 6);
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < 3; ++j) pw[(size_t)i * 3 + j] = (float)pts_world[(size_t)i * 3 + j];  // already float-valued in the reference
        var9_to_6(var9 + (size_t)i * 9, v6.data() + (size_t)i * 6);
    }
    IM_CUDA(cudaMemcpyAsync(h->sb.pw, pw.data(), pw.size() * 4, cudaMemcpyHostToDevice, h->stream));
    IM_CUDA(cudaMemcpyAsync(h->sb.var, v6.data(), v6.size() * 8, cudaMemcpyHostToDevice, h->stream));
    IM_CUDA(cudaStreamSynchronize(h->stream));
    h->sb.n = n;
    h->last_n = n;
    return IMMESH_OK;
}
int immesh_voxelmap_build_pv(immesh_lio_t* h, const double* pts_world, const double* var9, int n) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    int rc = upload_pv(h, pts_world, var9, n);
    if (rc) return rc;
    launch_grow(h, n, 3);
    IM_CUDA(cudaGetLastError());
    return check_flags(h);
}
int immesh_voxelmap_update_pv(immesh_lio_t* h, const double* pts_world, const double* var9, int n) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    int rc = upload_pv(h, pts_world, var9, n);
    if (rc) return rc;
    launch_grow(h, n, 2);
    IM_CUDA(cudaGetLastError());
    return check_flags(h);
}
int immesh_residual_build_pv(immesh_lio_t* h, const double* pts_body, const double* pts_world, const double* var9, int n, int* index_layer,
                             double* ptpl, int cap, int* n_out) {
    if (!h || !n_out || ((!pts_body || !pts_world || !var9) && n > 0)) return im_fail(IMMESH_E_INVALID, "null argument");
    if (n < 0 || n > h->max_scan) return im_fail(IMMESH_E_CAPACITY, "too many points");
    *n_out = 0;
    if (n == 0) return IMMESH_OK;
    std::vector<double> v6((size_t)n * 6);
    for (int i = 0; i < n; ++i) var9_to_6(var9 + (size_t)i * 9, v6.data() + (size_t)i * 6);
    IM_CUDA(cudaMemcpyAsync(h->sb.var, v6.data(), v6.size() * 8, cudaMemcpyHostToDevice, h->stream));
    IM_CUDA(cudaMemcpyAsync(h->sb.p_imu, pts_world, (size_t)n * 24, cudaMemcpyHostToDevice, h->stream));  // p_imu doubles as scratch for the world points
    IM_LAUNCH(k_match_pv, grid_for(h, n, 128), 128, 0, h->stream, h->map, h->P, h->sb, h->sb.p_imu, n);
    IM_CUDA(cudaGetLastError());
    std::vector<int> node(n), layer(n);
    IM_CUDA(cudaMemcpyAsync(node.data(), h->sb.match_node, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaMemcpyAsync(layer.data(), h->sb.match_layer, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaStreamSynchronize(h->stream));
    int m = 0, nn = 0;
    for (int i = 0; i < n; ++i) nn = node[i] > nn ? node[i] : nn;
    std::vector<PlaneRec> planes;
    // fetch only the matched plane records
    std::vector<int> uniq;
    for (int i = 0; i < n; ++i) if (node[i] >= 0) uniq.push_back(node[i]);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    planes.resize(uniq.size());
    for (size_t u = 0; u < uniq.size(); ++u)
        IM_CUDA(cudaMemcpyAsync(&planes[u], h->map.planes + uniq[u], sizeof(PlaneRec), cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n; ++i) {
        if (node[i] < 0) continue;
        if (m < cap) {
            const PlaneRec& pl = planes[std::lower_bound(uniq.begin(), uniq.end(), node[i]) - uniq.begin()];
            if (index_layer) { index_layer[2 * m] = i; index_layer[2 * m + 1] = layer[i]; }
            if (ptpl) {
                double* o = ptpl + (size_t)m * 31;
                for (int j = 0; j < 3; ++j) { o[j] = pts_body[(size_t)i * 3 + j]; o[3 + j] = pl.normal[j]; o[6 + j] = pl.center[j]; }
                o[9] = (double)pl.d;
                for (int j = 0; j < 21; ++j) o[10 + j] = pl.pv[j];
            }
        }
        ++m;
    }
    *n_out = m;
    return check_flags(h);
}

int immesh_lio_iter_stats(immesh_lio_t* h, int it, double* out) {
    if (!h || !out || it < 0 || it >= IM_MAX_ITER) return im_fail(IMMESH_E_INVALID, "bad argument");
    IterStats s;
    IM_CUDA(cudaMemcpyAsync(&s, &h->d_ctrl->stats[it], sizeof(IterStats), cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaStreamSynchronize(h->stream));
    std::memcpy(out, s.HTH, 36 * 8);
    std::memcpy(out + 36, s.HTz, 6 * 8);
    out[42] = s.n_match; out[43] = s.total_residual;
    std::memcpy(out + 44, s.solution, 18 * 8);
    out[62] = s.converged;
    return IMMESH_OK;
}

int immesh_lio_matches(immesh_lio_t* h, int* plane_layer, int n) {
    if (!h || !plane_layer || n > h->max_scan) return im_fail(IMMESH_E_INVALID, "bad argument");
    std::vector<int> node(n), layer(n);
    IM_CUDA(cudaMemcpyAsync(node.data(), h->sb.match_node, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaMemcpyAsync(layer.data(), h->sb.match_layer, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n; ++i) plane_layer[i] = node[i] >= 0 ? layer[i] : -1;
    return IMMESH_OK;
}

int immesh_lio_match_nodes(immesh_lio_t* h, int* node, int n) {
    if (!h || !node || n > h->max_scan) return im_fail(IMMESH_E_INVALID, "bad argument");
    IM_CUDA(cudaMemcpyAsync(node, h->sb.match_node, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaStreamSynchronize(h->stream));
    return IMMESH_OK;
}

int64_t immesh_voxelmap_dump(immesh_lio_t* h, double* rows, int64_t cap_rows) {
    if (!h) return -1;
    int counters[16];
    if (cudaMemcpy(counters, h->d_counters, sizeof(counters), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    const int nn = counters[0] < h->max_nodes ? counters[0] : h->max_nodes;
    std::vector<unsigned long long> keys(h->cap);
    std::vector<int> roots(h->cap);
    std::vector<NodeRec> nodes(nn > 0 ? nn : 1);
    std::vector<PlaneRec> planes(nn > 0 ? nn : 1);
    cudaMemcpy(keys.data(), h->map.keys, h->cap * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(roots.data(), h->map.root_node, h->cap * 4, cudaMemcpyDeviceToHost);
    if (nn > 0) {
        cudaMemcpy(nodes.data(), h->map.nodes, (size_t)nn * sizeof(NodeRec), cudaMemcpyDeviceToHost);
        cudaMemcpy(planes.data(), h->map.planes, (size_t)nn * sizeof(PlaneRec), cudaMemcpyDeviceToHost);
    }
    return dump_voxelmap(keys.data(), roots.data(), h->cap, nodes.data(), planes.data(), rows, cap_rows);
}

int immesh_voxelmap_counts(immesh_lio_t* h, int64_t* out) {
    if (!h || !out) return im_fail(IMMESH_E_INVALID, "null argument");
    int counters[16];
    IM_CUDA(cudaMemcpy(counters, h->d_counters, sizeof(counters), cudaMemcpyDeviceToHost));
    out[0] = counters[5]; out[1] = counters[0]; out[2] = counters[1]; out[3] = counters[4];
    return IMMESH_OK;
}

}  // extern "C"
