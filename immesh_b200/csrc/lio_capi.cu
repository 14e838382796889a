// immesh_b200 -- CUDA kernels (sm_100a) and C-ABI host orchestration of the localization path.
// Kernel bodies live in lio_core.cuh / voxelmap.cuh; this file adds the launch geometry, the
// block-level integer reduction of the normal equations, stream/event plumbing and the C ABI.
// There is no CPU path: every entry point needs a CUDA device and fails loudly without one.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
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
__global__ void __launch_bounds__(128) k_prepare(LioParams P, ScanBuf sb_) {
    const ScanBuf sb = scan_load_dyn(sb_);
    const int n = sb.n;
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

// K2+K3.  Default (one GPU): two kernels per IESKF iteration.
//   k_match : 8 lanes per scan point.  The walk over a root voxel's octree (build_single_residual visits EVERY plane below the root,
//             up to 1 + 8 + 64 for max_layer 2, each a dependent chain of ~150 double operations behind a record load) is split by
//             first-level child over the 8 lanes and merged with three shuffles, in the serial walk's order (match_in_voxel_lane).
//   k_terms : one thread per scan point: the 29 fixed-point normal-equation terms of the matched points; the block's sums are reduced
//             with warp shuffles (integer adds: exact, order-free) and folded into the iteration's global accumulators with 60
//             atomics; the last block to finish runs the 6x6 IESKF update (ieskf_solve).
// k_residual is the two steps in one thread per point (used by immesh_lio_residual_build; IMMESH_LIO_SPLIT=0 selects it everywhere).
#define RES_THREADS 128
#define MATCH_THREADS 128
#define MATCH_LANES 8
struct MatchCombine8 {
    unsigned int mask;   // the 8 lanes of this point inside the warp; 0: one lane per point
    __device__ void operator()(bool* ok, MatchResult* best) const {
        if (mask == 0u) return;   // one lane per point: nothing to merge
        const int lane = threadIdx.x & 31;
        *ok = (__ballot_sync(mask, *ok) & mask) != 0u;
        double prob = best->prob;
        int src = lane;
#pragma unroll
        for (int o = MATCH_LANES / 2; o > 0; o >>= 1) {
            const double op = __shfl_xor_sync(mask, prob, o);
            const int os = __shfl_xor_sync(mask, src, o);
            if (op > prob || (op == prob && os < src)) { prob = op; src = os; }   // first lane, in lane order, holding the largest probability
        }
        best->node = __shfl_sync(mask, best->node, src);
        best->layer = __shfl_sync(mask, best->layer, src);
        best->prob = prob;
    }
};
__global__ void __launch_bounds__(MATCH_THREADS, 8) k_match(VoxelMapDev map, LioParams P, ScanBuf sb_, LioCtrl* ctrl) {
    __shared__ double s_state[24 + 6 * 18];
    IM_STAMP(0, 0);
    // the stop flag, the scan's device-resident description and the state are loaded together (one trip to L2 instead of three)
    const int stop = ctrl->stop;
    const ScanBuf sb = scan_load_dyn(sb_);
    for (int i = threadIdx.x; i < 24 + 6 * 18; i += blockDim.x) s_state[i] = ctrl->state[i];
    const int n = sb.n;
    if (stop || n <= 0) return;
    IM_STAMP(1, n);
    // 8 lanes per point while the scan fits the machine about twice over (the grid is one resident wave); beyond that every SM has
    // more than enough points to hide the walk's latency and the lanes would only multiply the instruction count: one thread per point
    const long long threads = (long long)gridDim.x * blockDim.x;
    const int lanes = ((long long)n * MATCH_LANES <= 2 * threads) ? MATCH_LANES : 1;
    const int per_block = MATCH_THREADS / lanes;
    if ((long long)blockIdx.x * per_block >= n) return;   // block-uniform; the grid is sized for the machine, not for the scan
    __syncthreads();
    IM_STAMP(2, 0);
    const int sub = threadIdx.x & (lanes - 1);
    MatchCombine8 comb;
    comb.mask = (lanes == 1) ? 0u : 0xFFu << ((threadIdx.x & 31) & ~(MATCH_LANES - 1));
    for (int i = blockIdx.x * per_block + (threadIdx.x / lanes); i < n; i += gridDim.x * per_block)
        residual_match_lanes(map, P, sb, s_state, i, sub, lanes, comb);
    IM_STAMP(7, 0);
}

template <bool MATCHED>
__device__ __forceinline__ void residual_block(const VoxelMapDev& map, const LioParams& P, const ScanBuf& sb_, LioCtrl* ctrl, int iter, int fused_solve) {
    __shared__ double s_state[24 + 6 * 18];
    __shared__ long long s_part[RES_THREADS / 32][IM_NTERMS];
    IM_STAMP(10, 0);
    // the stop flag, the scan's device-resident description and the state -- rot/pos and the 6x6 pose block of the covariance, the only
    // parts this pass reads -- are loaded together (one trip to L2 instead of three)
    const int stop = ctrl->stop;
    const ScanBuf sb = scan_load_dyn(sb_);
    for (int i = threadIdx.x; i < 24 + 6 * 18; i += blockDim.x) s_state[i] = ctrl->state[i];
    const int n = sb.n;
    if (stop || n <= 0) return;   // empty scan: no iteration runs (iters_run stays 0), as with an empty map
    // the grid is sized for the largest scan (constant launch sequence); blocks without points leave at once and are not counted
    const int nb_active = min((int)gridDim.x, (n + RES_THREADS - 1) / RES_THREADS);
    if ((int)blockIdx.x >= nb_active) return;
    __syncthreads();
    IM_STAMP(11, 0);
    long long acc[IM_NTERMS];
#pragma unroll
    for (int k = 0; k < IM_NTERMS; ++k) acc[k] = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        long long t[IM_NTERMS];
        const bool hit = MATCHED ? residual_point_matched(map, P, sb, s_state, i, t, map.err) : residual_point(map, P, sb, s_state, i, t, map.err);
        if (hit) {
#pragma unroll
            for (int k = 0; k < IM_NTERMS; ++k) acc[k] += t[k];
        }
    }
    IM_STAMP(12, acc[0] + acc[27]);
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
    IM_STAMP(13, 0);
    if (!fused_solve) return;
    // the last block to publish its sums runs the IESKF update of this iteration (6x6 form, lio_core.cuh: ieskf_solve)
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ctrl->blocks_done[iter], 1) == nb_active - 1) ? 1 : 0;
    __syncthreads();
    IM_STAMP(14, s_last);
    if (s_last) {
        __shared__ SolveScratch S;
        __threadfence();
        ieskf_solve(P, ctrl, iter, &S, threadIdx.x, blockDim.x);
    }
}
__global__ void __launch_bounds__(RES_THREADS) k_residual(VoxelMapDev map, LioParams P, ScanBuf sb_, LioCtrl* ctrl, int iter, int fused_solve) {
    residual_block<false>(map, P, sb_, ctrl, iter, fused_solve);
}
__global__ void __launch_bounds__(RES_THREADS) k_terms(VoxelMapDev map, LioParams P, ScanBuf sb_, LioCtrl* ctrl, int iter, int fused_solve) {
    residual_block<true>(map, P, sb_, ctrl, iter, fused_solve);
}

// sharded VoxelMap: pass 1 (match where this rank owns the voxel, publish bits) and pass 2 (terms + integer reduction)
__global__ void __launch_bounds__(RES_THREADS) k_shard_pass1(VoxelMapDev map, LioParams P, ScanBuf sb_, LioCtrl* ctrl, unsigned int* bits, int words_cap) {
    __shared__ double s_state[24 + 6 * 18];
    if (ctrl->stop) return;
    const ScanBuf sb = scan_load_dyn(sb_);
    const int n = sb.n;
    if (n <= 0) return;
    const int words = words_cap;
    for (int i = threadIdx.x; i < 24 + 6 * 18; i += blockDim.x) s_state[i] = ctrl->state[i];
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) shard_pass1_point(map, P, sb, s_state, i, bits, bits + words);
}
__global__ void __launch_bounds__(RES_THREADS) k_shard_pass2(VoxelMapDev map, LioParams P, ScanBuf sb_, LioCtrl* ctrl, int iter, const unsigned int* bits, int words_cap) {
    __shared__ double s_state[24 + 6 * 18];
    __shared__ long long s_part[RES_THREADS / 32][IM_NTERMS];
    if (ctrl->stop) return;
    const ScanBuf sb = scan_load_dyn(sb_);
    const int n = sb.n;
    if (n <= 0) return;
    const int words = words_cap;
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
__global__ void __launch_bounds__(RES_THREADS) k_shard_pass1_p2p(VoxelMapDev map, LioParams P, ScanBuf sb_, LioCtrl* ctrl, LioPeers pe, int epoch_off) {
    __shared__ double s_state[24 + 6 * 18];
    __shared__ int s_last;
    if (ctrl->stop) return;
    const ScanBuf sb = scan_load_dyn(sb_);
    const int n = sb.n;
    if (n <= 0) return;
    const unsigned long long epoch = ctrl->dyn.epoch + (unsigned long long)epoch_off;
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
__global__ void __launch_bounds__(RES_THREADS) k_shard_pass2_p2p(VoxelMapDev map, LioParams P, ScanBuf sb_, LioCtrl* ctrl, int iter, LioPeers pe, int epoch_off) {
    __shared__ double s_state[24 + 6 * 18];
    __shared__ long long s_part[RES_THREADS / 32][IM_NTERMS];
    __shared__ int s_last;
    if (ctrl->stop) return;
    const ScanBuf sb = scan_load_dyn(sb_);
    const int n = sb.n;
    if (n <= 0) return;
    const unsigned long long epoch = ctrl->dyn.epoch + (unsigned long long)epoch_off;
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
#define SOLVE_THREADS 128
__global__ void __launch_bounds__(SOLVE_THREADS) k_solve_p2p(LioParams P, LioCtrl* ctrl, int iter, LioPeers pe, int epoch_off, int* err) {
    __shared__ SolveScratch S;
    if (ctrl->stop || (ctrl->dyn.n_dev ? *ctrl->dyn.n_dev : ctrl->dyn.n) <= 0) return;
    const unsigned long long epoch = ctrl->dyn.epoch + (unsigned long long)epoch_off;
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
    ieskf_solve(P, ctrl, iter, &S, threadIdx.x, blockDim.x);
}
// stand-alone solve (NCCL-sharded path: the sums are complete only after the all-reduce)
__global__ void __launch_bounds__(SOLVE_THREADS) k_solve(LioParams P, LioCtrl* ctrl, int iter) {
    __shared__ SolveScratch S;
    if (ctrl->stop || (ctrl->dyn.n_dev ? *ctrl->dyn.n_dev : ctrl->dyn.n) <= 0) return;
    ieskf_solve(P, ctrl, iter, &S, threadIdx.x, blockDim.x);
}

#define PREDICT_THREADS 352
__global__ void __launch_bounds__(PREDICT_THREADS) k_predict(LioCtrl* ctrl) {
    __shared__ double T[324], Fx[324];
    const double dt = ctrl->dyn.dt;
    if (!(dt > 0)) return;   // block-uniform
    predict_const_vel(ctrl->state, dt, ctrl->dyn.cov_gyr, ctrl->dyn.cov_acc, T, Fx, threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(128) k_grow_point(VoxelMapDev map, LioParams P, ScanBuf sb_, LioCtrl* ctrl, int mode) {
    const ScanBuf sb = scan_load_dyn(sb_);
    const int n = sb.n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) grow_point(map, P, sb, ctrl->state, i, mode);
}
__global__ void __launch_bounds__(128) k_grow_segment(ScanBuf sb) {
    const int nt = *sb.n_touched;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) grow_segment(sb, t);
}
__global__ void __launch_bounds__(128) k_grow_scatter(ScanBuf sb_) {
    const ScanBuf sb = scan_load_dyn(sb_);
    const int n = sb.n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) grow_scatter(sb, i);
}
// Most touched root voxels receive one or two points of a scan and stay in an append-only regime (a node still collecting its
// first points, or a planar node between two refits): no plane fit, no split, no descent.  k_grow_simple finishes those -- same
// appends, same order (ascending (var_contrast, index)), same arithmetic as update_octo_tree -- with a warp per voxel in a lean
// kernel (few registers: many more warps in flight than the general kernel below, which carries the plane fit and the octree
// split); only the other voxels go on to k_grow_voxel through a compact list.  At the 1M-point / 0.2 m configuration 175 k voxels
// are touched per scan and nearly all are of the first kind.  (A THREAD per simple voxel was measured and rejected: one thread's
// 60 dependent moment read-modify-writes take ~5x longer than the warp's two per lane, profiles/README.md.)
#define GROW_SIMPLE_MAX 4
__global__ void __launch_bounds__(128) k_grow_simple(VoxelMapDev map, LioParams P, ScanBuf sb, int mode, int* complex_list, int* n_complex) {
    const int lane = threadIdx.x & 31;
    const int nt = *sb.n_touched;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < nt; t += nwarps) {
        const int slot = sb.touched[t];
        const int cnt = sb.slot_count[slot], off = sb.slot_offset[slot];
        const int root = map.root_node[slot];
        int kind = 0;   // 0: general kernel, 1: append all, 2: drop all (frozen planar node)
        if (root >= 0 && (mode & 1) == 0 && cnt <= GROW_SIMPLE_MAX) {
            const NodeRec& n = map.nodes[root];
            if (!n.init_octo) kind = (n.n_pts + cnt <= P.layer_init[n.layer]) ? 1 : 0;
            else if (map.planes[root].is_plane) kind = !n.update_enable ? 2 : ((n.new_points + cnt <= 5 && n.n_pts + cnt < P.max_points) ? 1 : 0);
        }
        kind = __shfl_sync(0xffffffffu, kind, 0);
        if (kind == 0) {
            if (lane == 0) complex_list[atomicAdd(n_complex, 1)] = t;
            continue;
        }
        if (kind == 1) {
            int idx[GROW_SIMPLE_MAX];
            double key[GROW_SIMPLE_MAX];
#pragma unroll
            for (int a = 0; a < GROW_SIMPLE_MAX; ++a) {
                idx[a] = 0x7fffffff; key[a] = 0.0;
                if (a < cnt) { idx[a] = sb.seg[off + a]; key[a] = sb.sortkey[idx[a]]; }
            }
#pragma unroll
            for (int a = 1; a < GROW_SIMPLE_MAX; ++a)       // insertion sort on (key, index): the order std::sort(var_contrast) with ties by index gives
#pragma unroll
                for (int b = a; b > 0; --b) {
                    const bool lt = b < cnt && (key[b] < key[b - 1] || (key[b] == key[b - 1] && idx[b] < idx[b - 1]));
                    if (lt) { const int ti = idx[b]; idx[b] = idx[b - 1]; idx[b - 1] = ti; const double tk = key[b]; key[b] = key[b - 1]; key[b - 1] = tk; }
                }
#pragma unroll
            for (int a = 0; a < GROW_SIMPLE_MAX; ++a) {
                if (a >= cnt) break;
                const int i = idx[a];
                const float px = sb.pw[(size_t)i * 3 + 0], py = sb.pw[(size_t)i * 3 + 1], pz = sb.pw[(size_t)i * 3 + 2];
                if (lane == 0) {
                    map.nodes[root].new_points += 1;
                    node_append(map, root, px, py, pz, sb.var + (size_t)i * 6);
                }
                node_moments_add(map, root, px, py, pz, sb.var + (size_t)i * 6, lane, 32);
                __syncwarp();
            }
        }
        if (lane == 0) {
            sb.slot_count[slot] = 0;
            sb.slot_cursor[slot] = 0;
        }
    }
}
// one warp per remaining touched root voxel, voxels claimed dynamically (their cost varies by orders of magnitude)
__global__ void __launch_bounds__(128) k_grow_voxel(VoxelMapDev map, LioParams P, ScanBuf sb, int mode, int* sorted_scratch, int* work_counter, const int* complex_list, const int* n_complex) {
    const int lane = threadIdx.x & 31;
    const int nt = *n_complex;
    while (true) {
        int t = 0;
        if (lane == 0) t = atomicAdd(work_counter, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= nt) break;
        grow_voxel(map, P, sb, complex_list[t], mode, lane, 32, sorted_scratch);
    }
}
// last kernel of a scan: recycle chunks, reset the per-scan counters, publish the converged pose for the mesher's frame (which
// runs behind on another stream) and assemble the block the host reads back with one D2H copy
__global__ void k_grow_finish(VoxelMapDev map, ScanBuf sb, int* work_counter, LioCtrl* ctrl, const int* counters) {
    recycle_chunks(map, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) {
        *sb.n_touched = 0;
        *sb.seg_top = 0;
        *work_counter = 0;
        work_counter[3] = 0;   // n_complex (d_counters + 11)
    }
    if (ctrl) {
        const int si = ctrl->dyn.scan_idx;
        for (int i = threadIdx.x; i < IM_STATE_DOUBLES; i += blockDim.x) ctrl->out.state[i] = ctrl->state[i];
        if (threadIdx.x < 12) ctrl->pose_ring[si & (IM_POSE_RING - 1)][threadIdx.x] = ctrl->state[threadIdx.x];
        if (threadIdx.x < 16) ctrl->out.counters[threadIdx.x] = counters[threadIdx.x];
        if (threadIdx.x == 0) { ctrl->out.iters_run = ctrl->iters_run; ctrl->out.scan_idx = si; }
    }
    __syncthreads();
    if (threadIdx.x == 0 && map.stat) { map.stat[0] = 0; map.stat[1] = 0; }   // per-scan work counters (reported through LioOut::counters[9..10])
}
// BuildResidualListOMP on caller-supplied Point_with_var data (world point + covariance per point)
__global__ void __launch_bounds__(128) k_match_pv(VoxelMapDev map, LioParams P, ScanBuf sb, const double* pw, int n) {   // API call: explicit n
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

// Grids are fixed per handle (sized for max_scan_points, at most 8 waves of blocks; every per-point kernel is a grid-stride
// loop over the n it reads from the device-resident ScanDyn), so that the captured launch sequence is identical for every scan.
static int grid_fixed(const immesh_lio* h, int threads, int max_waves = 8) {
    static const int by_n = std::getenv("IMMESH_DEBUG_GRID_N") ? std::atoi(std::getenv("IMMESH_DEBUG_GRID_N")) : 0;   // experiments: size the grids by the scan
    long long g = ((long long)(by_n ? (h->last_n > 0 ? h->last_n : 1) : h->max_scan) + threads - 1) / threads;
    const long long cap = (long long)h->n_sm * max_waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

#if defined(IM_DEBUG_STAMPS)
extern "C" int immesh_debug_stamps(long long* out64) {   // debug variant only (tools/debug/build_stamps.sh); not declared in include/
    return cudaMemcpyFromSymbol(out64, immesh::g_stamps, 64 * sizeof(long long)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" {

int immesh_lio_create(const immesh_lio_config* cfg, immesh_lio_t** out) {
    if (!cfg || !out) return im_fail(IMMESH_E_INVALID, "null argument");
    if (cfg->max_iteration < 1 || cfg->max_iteration > IM_MAX_ITER) return im_fail(IMMESH_E_INVALID, "max_iteration must be in [1,8]");
    if (cfg->max_layer < 0 || cfg->max_layer > 4) return im_fail(IMMESH_E_INVALID, "max_layer must be in [0,4]");
    if (!(cfg->voxel_size > 0) || cfg->max_points_size < 1) return im_fail(IMMESH_E_INVALID, "voxel_size and max_points_size must be positive");
    for (int i = 0; i < 5; ++i)
        if (cfg->layer_init_size[i] < 0) return im_fail(IMMESH_E_INVALID, "layer_init_size must be non-negative");
    if (cfg->hash_capacity_log2 != 0 && (cfg->hash_capacity_log2 < 10 || cfg->hash_capacity_log2 > 30)) return im_fail(IMMESH_E_INVALID, "hash_capacity_log2 must be 0 (default) or in [10,30]");
    if (cfg->max_nodes < 0 || cfg->max_chunks < 0 || cfg->max_scan_points < 0) return im_fail(IMMESH_E_INVALID, "capacities must be non-negative (0 = default)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return im_fail(IMMESH_E_NO_DEVICE, "no CUDA device: immesh_b200 has no CPU path");
    immesh_lio* h = new immesh_lio();
    // every failure below releases what has been created so far (immesh_lio_destroy copes with a partially built handle)
#define IM_CREATE(expr)                                                                     \
    do {                                                                                    \
        cudaError_t im_e_ = (expr);                                                         \
        if (im_e_ != cudaSuccess) { immesh_lio_destroy(h); return immesh::im_fail_cuda(im_e_, __FILE__, __LINE__); } \
    } while (0)
    fill_params(cfg, h->P);
    h->bps = std::getenv("IMMESH_LIO_BPS") ? std::atoi(std::getenv("IMMESH_LIO_BPS")) : 4;
    h->use_graph = std::getenv("IMMESH_GRAPH") ? std::atoi(std::getenv("IMMESH_GRAPH")) : 1;
    const int caplog = cfg->hash_capacity_log2 ? cfg->hash_capacity_log2 : 22;
    h->cap = (size_t)1 << caplog;
    h->max_nodes = cfg->max_nodes ? cfg->max_nodes : (4 << 20);
    h->max_chunks = cfg->max_chunks ? cfg->max_chunks : (4 << 20);
    h->max_scan = cfg->max_scan_points ? cfg->max_scan_points : (2 << 20);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, dev);
    IM_CREATE(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, im_stream_priority("IMMESH_LIO_PRIO")));
    for (auto& e : h->ev) IM_CREATE(cudaEventCreate(&e));
    for (int i = 0; i < IM_SLOTS; ++i) IM_CREATE(cudaEventCreateWithFlags(&h->ev_slot[i], cudaEventDisableTiming));
    IM_CREATE(cudaEventCreateWithFlags(&h->ev_pose, cudaEventDisableTiming));
    VoxelMapDev& m = h->map;
    IM_CREATE(dev_alloc(h, &m.keys, h->cap, 0xFF));
    IM_CREATE(dev_alloc(h, &m.root_node, h->cap, 0xFF));
    m.cap_mask = (unsigned)(h->cap - 1);
    IM_CREATE(dev_alloc(h, &m.nodes, (size_t)h->max_nodes));
    IM_CREATE(dev_alloc(h, &m.planes, (size_t)h->max_nodes));
    IM_CREATE(dev_alloc(h, &m.chunks, (size_t)h->max_chunks));
    IM_CREATE(dev_alloc(h, &m.avail, (size_t)h->max_chunks));
    IM_CREATE(dev_alloc(h, &m.pending, (size_t)h->max_chunks));
    IM_CREATE(dev_alloc(h, &h->d_counters, 16, 0));
    m.node_count = h->d_counters + 0; m.chunk_bump = h->d_counters + 1; m.avail_top = h->d_counters + 2; m.pending_n = h->d_counters + 3;
    m.err = h->d_counters + 4; m.n_roots = h->d_counters + 5; m.stat = h->d_counters + 9;
    m.max_nodes = h->max_nodes; m.max_chunks = h->max_chunks;
    ScanBuf& sb = h->sb;
    const size_t ms = (size_t)h->max_scan;
    float* d_body = nullptr;
    IM_CREATE(dev_alloc(h, &d_body, IM_SLOTS * ms * 3));
    IM_CREATE(cudaStreamCreateWithFlags(&h->stream_up, cudaStreamNonBlocking));
    for (int i = 0; i < IM_SLOTS; ++i) IM_CREATE(cudaEventCreateWithFlags(&h->ev_up[i], cudaEventDisableTiming));
    sb.body = d_body;
    h->d_body_own = d_body;
    IM_CREATE(dev_alloc(h, &sb.body_cov, ms * 6));
    IM_CREATE(dev_alloc(h, &sb.p_imu, ms * 3));
    IM_CREATE(dev_alloc(h, &sb.bv_imu, ms * 6));
    IM_CREATE(dev_alloc(h, &sb.match_node, ms, 0xFF));
    IM_CREATE(dev_alloc(h, &sb.match_layer, ms, 0));
    IM_CREATE(dev_alloc(h, &sb.pw, ms * 3));
    IM_CREATE(dev_alloc(h, &sb.var, ms * 6));
    IM_CREATE(dev_alloc(h, &sb.sortkey, ms));
    IM_CREATE(dev_alloc(h, &sb.slot, ms));
    IM_CREATE(dev_alloc(h, &sb.seg, ms));
    IM_CREATE(dev_alloc(h, &h->d_sorted, ms));
    IM_CREATE(dev_alloc(h, &h->d_complex, ms));
    IM_CREATE(dev_alloc(h, &sb.touched, ms));
    IM_CREATE(dev_alloc(h, &sb.slot_count, h->cap, 0));
    IM_CREATE(dev_alloc(h, &sb.slot_offset, h->cap, 0));
    IM_CREATE(dev_alloc(h, &sb.slot_cursor, h->cap, 0));
    sb.n_touched = h->d_counters + 6; sb.seg_top = h->d_counters + 7;
    sb.n = 0;
    IM_CREATE(dev_alloc(h, &h->d_ctrl, 1, 0));
    sb.dyn = &h->d_ctrl->dyn;
    IM_CREATE(cudaMallocHost((void**)&h->h_body, IM_SLOTS * ms * 3 * sizeof(float)));
    IM_CREATE(cudaMallocHost((void**)&h->h_state, (IM_STATE_DOUBLES + 64) * sizeof(double)));
    IM_CREATE(cudaMallocHost((void**)&h->h_ints, 64 * sizeof(int)));
    IM_CREATE(cudaMallocHost((void**)&h->h_dyn, IM_SLOTS * sizeof(ScanDyn)));
    IM_CREATE(cudaMallocHost((void**)&h->h_out, IM_SLOTS * sizeof(LioOut)));
    std::memset(h->h_out, 0, IM_SLOTS * sizeof(LioOut));
    // StatesGroup(): identity rotation, cov = INIT_COV * I  (include/common_lib.h:201-211)
    std::memset(h->h_state, 0, IM_STATE_DOUBLES * sizeof(double));
    h->h_state[0] = h->h_state[4] = h->h_state[8] = 1.0;
    for (int i = 0; i < 18; ++i) h->h_state[24 + i * 18 + i] = 0.0000001;
    IM_CREATE(cudaMemcpy(h->d_ctrl->state, h->h_state, IM_STATE_DOUBLES * sizeof(double), cudaMemcpyHostToDevice));
    IM_CREATE(cudaDeviceSynchronize());
#undef IM_CREATE
    *out = h;
    return IMMESH_OK;
}

int immesh_lio_destroy(immesh_lio_t* h) {
    if (!h) return IMMESH_OK;
    if (h->stream) cudaStreamSynchronize(h->stream);
    h->graph.destroy();
    if (h->win.local) immesh::peer_window_close(h->win);
    if (h->nccl_comm && nccl().CommDestroy) nccl().CommDestroy(h->nccl_comm);
    for (void* p : h->allocs) cudaFree(p);
    if (h->h_body) cudaFreeHost(h->h_body);
    if (h->h_state) cudaFreeHost(h->h_state);
    if (h->h_ints) cudaFreeHost(h->h_ints);
    if (h->h_dyn) cudaFreeHost(h->h_dyn);
    if (h->h_out) cudaFreeHost(h->h_out);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->ev_pose) cudaEventDestroy(h->ev_pose);
    for (int i = 0; i < IM_SLOTS; ++i) if (h->ev_up[i]) cudaEventDestroy(h->ev_up[i]);
    if (h->stream_up) cudaStreamDestroy(h->stream_up);
    if (h->ev_mark) cudaEventDestroy(h->ev_mark);
    if (h->stream) cudaStreamDestroy(h->stream);
    for (int i = 0; i < IM_SLOTS; ++i) if (h->ev_slot[i]) cudaEventDestroy(h->ev_slot[i]);
    cudaGetLastError();
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
    if (nranks > IM_MAX_RANKS) return im_fail(IMMESH_E_INVALID, "at most 8 ranks (one node)");
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
    h->graph.destroy();   // the launch sequence changes
    h->words_cap = h->max_scan / 32 + 2;
    if (!h->d_bits) IM_CUDA(dev_alloc(h, &h->d_bits, (size_t)2 * h->words_cap, 0));
    // peer window over NVLink (CUDA IPC): the exchanges of the residual pass are then fused into its kernels.  IMMESH_SHARD_NCCL=1
    // keeps the NCCL all-reduce sequence (the baseline the fused path is measured against; same setting on every rank).  The
    // open is collective and its outcome is the same on every rank (peer_win.cuh).
    const char* force_nccl = std::getenv("IMMESH_SHARD_NCCL");
    if (!(force_nccl && force_nccl[0] == '1')) {
        const size_t bytes = IM_LIOWIN_BITS_OFF + (size_t)IM_MAX_RANKS * 2 * h->words_cap * sizeof(unsigned int);
        const cudaError_t e = immesh::peer_window_open(h->win, bytes, rank, nranks, comm, h->stream);
        if (e != cudaSuccess) {
            std::fprintf(stderr, "[immesh_b200] rank %d: peer window unavailable on some rank (%s here); all ranks use NCCL all-reduces for the sharded residual pass\n", rank, cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    return IMMESH_OK;
}
int immesh_lio_shard_transport(immesh_lio_t* h) {   // 0 = not sharded, 1 = NCCL all-reduces, 2 = fused peer-window exchange
    if (!h) return 0;
    if (h->P.shard_n <= 1) return 0;
    return h->win.ok ? 2 : 1;
}

int immesh_lio_set_state(immesh_lio_t* h, const double* s) {
    if (!h || !s) return im_fail(IMMESH_E_INVALID, "null argument");
    IM_CUDA(cudaStreamSynchronize(h->stream));
    std::memcpy(h->h_state, s, IM_STATE_DOUBLES * sizeof(double));
    IM_CUDA(cudaMemcpyAsync(h->d_ctrl->state, h->h_state, IM_STATE_DOUBLES * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    IM_CUDA(cudaStreamSynchronize(h->stream));
    h->pose_pub_idx = -1;
    return IMMESH_OK;
}
int immesh_lio_get_state(immesh_lio_t* h, double* s) {
    if (!h || !s) return im_fail(IMMESH_E_INVALID, "null argument");
    IM_CUDA(cudaMemcpyAsync(h->h_state, h->d_ctrl->state, IM_STATE_DOUBLES * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaStreamSynchronize(h->stream));
    std::memcpy(s, h->h_state, IM_STATE_DOUBLES * sizeof(double));
    return IMMESH_OK;
}

// true when `p` is page-locked host memory the device can read directly (cudaMallocHost / cudaHostRegister): such a scan is
// copied straight from the caller's buffer, without the staging memcpy
static bool host_ptr_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
static int push_dyn(immesh_lio* h, int n, int slot = 0, double dt = 0.0, double cov_gyr = 0.0, double cov_acc = 0.0);
// per-scan inputs: the scan itself (host: staged through pinned slot `slot` unless the caller's buffer is pinned already) and
// the ScanDyn block every kernel of the sequence reads; both in stream order, so the previous scan is never disturbed
static int upload_scan(immesh_lio* h, const float* body, int n, int on_device = 0, int slot = 0, double dt = 0.0, double cov_gyr = 0.0, double cov_acc = 0.0) {
    if ((!body && n > 0) || n < 0) return im_fail(IMMESH_E_INVALID, "bad scan");
    if (n > h->max_scan) return im_fail(IMMESH_E_CAPACITY, "scan larger than max_scan_points");
    if (on_device) {
        h->sb.body = body;  // caller-owned device buffer, must stay valid until the next call on this handle
    } else {
        float* dst = h->d_body_own + (size_t)slot * h->max_scan * 3;   // the slot's previous user has finished (ev_slot waited on by the caller)
        h->sb.body = dst;
        if (n > 0) {
            const float* src = body;
            if (!host_ptr_is_pinned(body)) {
                float* stage = h->h_body + (size_t)slot * h->max_scan * 3;
                std::memcpy(stage, body, (size_t)n * 3 * sizeof(float));
                src = stage;
            }
            // on the upload stream, so that the copy overlaps the kernels of the scan before; the compute stream joins it
            IM_CUDA(cudaMemcpyAsync((void*)dst, src, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream_up));
            IM_CUDA(cudaEventRecord(h->ev_up[slot], h->stream_up));
            IM_CUDA(cudaStreamWaitEvent(h->stream, h->ev_up[slot], 0));
        }
    }
    return push_dyn(h, n, slot, dt, cov_gyr, cov_acc);
}
static int push_dyn(immesh_lio* h, int n, int slot, double dt, double cov_gyr, double cov_acc) {
    h->sb.n = n;
    h->last_n = n;
    ScanDyn& d = h->h_dyn[slot];
    d.body = h->sb.body;
    d.n = n;
    d.scan_idx = h->scan_counter;
    d.dt = dt; d.cov_gyr = cov_gyr; d.cov_acc = cov_acc;
    d.epoch = h->win.epoch;
    d.mode = 0; d.pad_ = 0;
    d.n_dev = h->next_n_dev;
    h->next_n_dev = nullptr;
    h->dyn_last = d;
    h->dyn_last.dt = 0.0;
    IM_CUDA(cudaMemcpyAsync(&h->d_ctrl->dyn, &d, sizeof(ScanDyn), cudaMemcpyHostToDevice, h->stream));
    return IMMESH_OK;
}
static int lio_flags_status(int err) {
    if (err & (IM_ERR_NODE_POOL | IM_ERR_CHUNK_POOL | IM_ERR_HASH_FULL | IM_ERR_SEG_POOL)) return im_fail(IMMESH_E_CAPACITY, "device pool overflow (raise the capacities in immesh_lio_config)");
    if (err & (IM_ERR_KEY_RANGE | IM_ERR_FX_RANGE)) return im_fail(IMMESH_E_RANGE, "coordinate / normal-equation term outside the representable range");
    if (err & IM_ERR_PEER_TIMEOUT) return im_fail(IMMESH_E_CUDA, "sharded mode: a peer rank did not publish its data in time (peer window epoch flag)");
    return IMMESH_OK;
}
static int check_flags(immesh_lio* h) {
    IM_CUDA(cudaMemcpyAsync(h->h_ints, h->d_counters, 16 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    IM_CUDA(cudaStreamSynchronize(h->stream));
    return lio_flags_status(h->h_ints[4]);
}
// map growth of the scan described by the device-resident ScanDyn
static void launch_grow(immesh_lio* h, int mode) {
    const int g = grid_fixed(h, 128);
    IM_LAUNCH(k_grow_point, g, 128, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, mode);
    IM_LAUNCH(k_grow_segment, grid_fixed(h, 128, 2), 128, 0, h->stream, h->sb);
    IM_LAUNCH(k_grow_scatter, g, 128, 0, h->stream, h->sb);
    IM_LAUNCH(k_grow_simple, h->n_sm * 16, 128, 0, h->stream, h->map, h->P, h->sb, mode, h->d_complex, h->d_counters + 11);
    IM_LAUNCH(k_grow_voxel, h->n_sm * h->bps, 128, 0, h->stream, h->map, h->P, h->sb, mode, h->d_sorted, h->d_counters + 8, (const int*)h->d_complex, (const int*)(h->d_counters + 11));
    IM_LAUNCH(k_grow_finish, 1, 256, 0, h->stream, h->map, h->sb, h->d_counters + 8, h->d_ctrl, (const int*)h->d_counters);
}
// IESKF iterations.  One GPU: per iteration ONE kernel -- the residual pass; its last block runs the 6x6 IESKF update.
// Sharded over peer windows: pass1(+publish bits) -> pass2(wait bits, +publish sums) -> solve(wait sums), exchanges fused in.
// Sharded over NCCL (baseline): pass1 -> all-reduce(bit words) -> pass2 -> all-reduce(58 int64 sums) -> solve.
static int launch_estimate(immesh_lio* h) {
    IM_LAUNCH(k_reset_scan, 2, 256, 0, h->stream, h->sb, h->d_ctrl, 1);
    IM_LAUNCH(k_prepare, grid_fixed(h, 128), 128, 0, h->stream, h->P, h->sb);
    const int g = grid_fixed(h, RES_THREADS, 4);
    if (h->P.shard_n > 1 && h->win.ok) {
        LioPeers pe;
        for (int r = 0; r < IM_MAX_RANKS; ++r) pe.w[r] = h->win.peer[r];
        pe.rank = h->win.rank; pe.n = h->win.n; pe.words_cap = h->words_cap; pe.pad = 0;
        for (int it = 0; it < h->P.max_iter; ++it) {
            // one epoch per iteration on top of the scan's base (ScanDyn::epoch, advanced once per scan); the two phases have separate flag rows
            IM_LAUNCH(k_shard_pass1_p2p, g, RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, pe, 1 + it);
            IM_LAUNCH(k_shard_pass2_p2p, g, RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, it, pe, 1 + it);
            IM_LAUNCH(k_solve_p2p, 1, SOLVE_THREADS, 0, h->stream, h->P, h->d_ctrl, it, pe, 1 + it, h->map.err);
        }
        return IMMESH_OK;
    }
    if (h->P.shard_n > 1) {
        const int words = h->words_cap;
        for (int it = 0; it < h->P.max_iter; ++it) {
            IM_CUDA(cudaMemsetAsync(h->d_bits, 0, (size_t)2 * words * sizeof(unsigned int), h->stream));
            IM_LAUNCH(k_shard_pass1, g, RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, h->d_bits, words);
            if (nccl().AllReduce(h->d_bits, h->d_bits, (size_t)2 * words, kNcclUint32, kNcclSum, h->nccl_comm, h->stream)) return im_fail(IMMESH_E_CUDA, "ncclAllReduce(bits) failed");
            IM_LAUNCH(k_shard_pass2, g, RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, it, (const unsigned int*)h->d_bits, words);
            if (nccl().AllReduce(&h->d_ctrl->acc[it][0], &h->d_ctrl->acc[it][0], (size_t)IM_NTERMS * 2, kNcclUint64, kNcclSum, h->nccl_comm, h->stream)) return im_fail(IMMESH_E_CUDA, "ncclAllReduce(acc) failed");
            IM_LAUNCH(k_solve, 1, SOLVE_THREADS, 0, h->stream, h->P, h->d_ctrl, it);
        }
        return IMMESH_OK;
    }
    static const int split = std::getenv("IMMESH_LIO_SPLIT") ? std::atoi(std::getenv("IMMESH_LIO_SPLIT")) : 1;
    if (!split) {
        for (int it = 0; it < h->P.max_iter; ++it) IM_LAUNCH(k_residual, g, RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, it, 1);
        return IMMESH_OK;
    }
    const int gm = h->n_sm * 8;   // one resident wave of k_match blocks (__launch_bounds__(128, 8)); k_match picks the lanes per point from n
    for (int it = 0; it < h->P.max_iter; ++it) {
        IM_LAUNCH(k_match, gm, MATCH_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl);
        IM_LAUNCH(k_terms, g, RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, it, 1);
    }
    return IMMESH_OK;
}

int immesh_lio_predict(immesh_lio_t* h, double dt, double cov_gyr, double cov_acc) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    IM_CUDA(cudaStreamSynchronize(h->stream));
    ScanDyn& d = h->h_dyn[0];   // the scan described by the block stays what it was; only the prediction inputs change
    d = h->dyn_last;
    d.dt = dt; d.cov_gyr = cov_gyr; d.cov_acc = cov_acc;
    IM_CUDA(cudaMemcpyAsync(&h->d_ctrl->dyn, &d, sizeof(ScanDyn), cudaMemcpyHostToDevice, h->stream));
    IM_LAUNCH(k_predict, 1, PREDICT_THREADS, 0, h->stream, h->d_ctrl);
    IM_CUDA(cudaGetLastError());
    IM_CUDA(cudaStreamSynchronize(h->stream));
    h->pose_pub_idx = -1;
    return IMMESH_OK;
}

int immesh_voxelmap_build(immesh_lio_t* h, const float* body, int n) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    IM_CUDA(cudaStreamSynchronize(h->stream));
    int rc = upload_scan(h, body, n);
    if (rc) return rc;
    launch_grow(h, 1);
    IM_CUDA(cudaGetLastError());
    return check_flags(h);
}

int immesh_lio_estimate(immesh_lio_t* h, const float* body, int n, int* iters_run) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    IM_CUDA(cudaStreamSynchronize(h->stream));
    h->win.epoch += IM_MAX_ITER + 1;   // epoch base of this scan (same sequence of calls on every rank)
    int rc = upload_scan(h, body, n);
    if (rc) return rc;
    rc = launch_estimate(h);
    if (rc) return rc;
    IM_CUDA(cudaGetLastError());
    IM_CUDA(cudaMemcpyAsync(h->h_ints + 32, &h->d_ctrl->iters_run, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    rc = check_flags(h);
    if (iters_run) *iters_run = h->h_ints[32];
    h->pose_pub_idx = -1;
    return rc;
}

int immesh_voxelmap_update(immesh_lio_t* h) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    launch_grow(h, 0);
    IM_CUDA(cudaGetLastError());
    return check_flags(h);
}

// queue predict + estimate + update for one scan; no host synchronisation unless both staging slots are busy.  Host work per
// scan in the pipelined form: (memcpy into the pinned slot unless the caller's buffer is pinned) + H2D scan + H2D ScanDyn +
// ONE cudaGraphLaunch (no node is touched) + ONE D2H of the LioOut block + one event.
static int lio_enqueue(immesh_lio_t* h, const float* body, int n, int on_device, double dt, double cov_gyr, double cov_acc, bool allow_graph) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    const int s = (++h->step_counter) & (IM_SLOTS - 1);
    if (h->slot_busy[s]) {
        const auto t0 = std::chrono::steady_clock::now();
        IM_CUDA(cudaEventSynchronize(h->ev_slot[s]));
        h->host_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        h->slot_busy[s] = 0;
    }
    h->win.epoch += IM_MAX_ITER + 1;   // epoch base of this scan: advanced exactly once per scan
    ++h->scan_counter;
    const bool timing = !(allow_graph && h->use_graph && !profiler().enabled && (h->P.shard_n <= 1 || h->win.ok));
    if (timing) IM_CUDA(cudaEventRecord(h->ev[0], h->stream));
    int rc = upload_scan(h, body, n, on_device, s, dt, cov_gyr, cov_acc);
    if (rc) return rc;
    // The pipelined entry points replay the scan's launch sequence as one CUDA graph, captured once: every per-scan value is
    // read from the device-resident ScanDyn, so a replay is a bare cudaGraphLaunch.  The blocking entry points launch
    // directly, with stage timing events in between.
    bool queued = false;
    int lrc = IMMESH_OK;
    if (!timing) {
        const unsigned sig = 1u | (h->win.ok ? 8u : 0u) | ((unsigned)h->P.max_iter << 8) | ((unsigned)h->P.shard_n << 16);
        queued = immesh::run_graphed_static(h->graph, sig, h->stream, [&] {
            IM_LAUNCH(k_predict, 1, PREDICT_THREADS, 0, h->stream, h->d_ctrl);
            lrc = launch_estimate(h);
            launch_grow(h, 0);
        }) == cudaSuccess;
        if (!queued) h->use_graph = 0;   // not expected; keep working through direct launches
    }
    if (!queued) {
        if (!timing) IM_CUDA(cudaEventRecord(h->ev[0], h->stream));
        IM_LAUNCH(k_predict, 1, PREDICT_THREADS, 0, h->stream, h->d_ctrl);
        IM_CUDA(cudaEventRecord(h->ev[1], h->stream));
        lrc = launch_estimate(h);
        IM_CUDA(cudaEventRecord(h->ev[2], h->stream));
        launch_grow(h, 0);
    }
    if (lrc) return lrc;
    IM_CUDA(cudaGetLastError());
    IM_CUDA(cudaMemcpyAsync(h->h_out + s, &h->d_ctrl->out, sizeof(LioOut), cudaMemcpyDeviceToHost, h->stream));
    if (!queued) IM_CUDA(cudaEventRecord(h->ev[3], h->stream));
    IM_CUDA(cudaEventRecord(h->ev_slot[s], h->stream));
    h->slot_busy[s] = 1;
    h->timed_last = queued ? 0 : 1;
    h->pose_pub_idx = h->scan_counter;
    return IMMESH_OK;
}
static int lio_wait_impl(immesh_lio_t* h, double* state_out, int* iters_run, bool timings) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    const int s = h->step_counter & (IM_SLOTS - 1);
    IM_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < IM_SLOTS; ++i) h->slot_busy[i] = 0;
    if (profiler().enabled) profiler().collect();
    const LioOut& o = h->h_out[s];
    if (state_out) std::memcpy(state_out, o.state, IM_STATE_DOUBLES * sizeof(double));
    if (iters_run) *iters_run = o.iters_run;
    if (timings && h->timed_last) {
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, h->ev[0], h->ev[3]);
        cudaEventElapsedTime(&b, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&c, h->ev[2], h->ev[3]);
        h->last_ms[0] = a; h->last_ms[1] = b; h->last_ms[2] = c;
    }
    return lio_flags_status(o.counters[4]);
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
// the scan AND its point count are device resident (e.g. produced by the device front-end): n is read at execution time
int immesh_lio_step_async_dev_n(immesh_lio_t* h, const float* d_body_xyz, int n_max, const int* d_n, double dt, double cov_gyr, double cov_acc) {
    if (!h || !d_n || n_max < 0) return im_fail(IMMESH_E_INVALID, "bad argument");
    h->next_n_dev = d_n;
    return lio_enqueue(h, d_body_xyz, n_max, 1, dt, cov_gyr, cov_acc, true);
}
int immesh_lio_wait(immesh_lio_t* h, double* state_out, int* iters_run) { return lio_wait_impl(h, state_out, iters_run, false); }
// queue a write of `bytes` bytes over a caller-provided device buffer on the localization stream (benchmark L2 flush)
int immesh_lio_enqueue_memset(immesh_lio_t* h, void* d_buf, size_t bytes) {
    if (!h || !d_buf) return im_fail(IMMESH_E_INVALID, "null argument");
    IM_CUDA(cudaMemsetAsync(d_buf, 0, bytes, h->stream));
    return IMMESH_OK;
}

// work counters of the last step (roofline byte model): [n, touched root voxels are not kept, plane refits, points read by the refits]
int immesh_lio_work_stats(immesh_lio_t* h, int64_t* out /*[4]*/) {
    if (!h || !out) return im_fail(IMMESH_E_INVALID, "null argument");
    const LioOut& o = h->h_out[h->step_counter & (IM_SLOTS - 1)];
    out[0] = h->last_n; out[1] = o.counters[5]; out[2] = o.counters[9]; out[3] = o.counters[10];
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
    IM_LAUNCH(k_prepare, grid_fixed(h, 128), 128, 0, h->stream, h->P, h->sb);
    IM_LAUNCH(k_residual, grid_fixed(h, RES_THREADS, 4), RES_THREADS, 0, h->stream, h->map, h->P, h->sb, h->d_ctrl, 0, 0);
    IM_LAUNCH(k_gather_ptpl, grid_fixed(h, 128), 128, 0, h->stream, h->map, h->sb, n, h->d_ptpl);
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
    return push_dyn(h, n);
}
int immesh_voxelmap_build_pv(immesh_lio_t* h, const double* pts_world, const double* var9, int n) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    int rc = upload_pv(h, pts_world, var9, n);
    if (rc) return rc;
    launch_grow(h, 3);
    IM_CUDA(cudaGetLastError());
    return check_flags(h);
}
int immesh_voxelmap_update_pv(immesh_lio_t* h, const double* pts_world, const double* var9, int n) {
    if (!h) return im_fail(IMMESH_E_INVALID, "null handle");
    int rc = upload_pv(h, pts_world, var9, n);
    if (rc) return rc;
    launch_grow(h, 2);
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
    IM_LAUNCH(k_match_pv, grid_fixed(h, 128), 128, 0, h->stream, h->map, h->P, h->sb, h->sb.p_imu, n);
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
