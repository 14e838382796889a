// immesh_b200 -- front-end in front of the hot path: pcl::VoxelGrid centroid down-sampling of a scan on the
// device, replacing  m_downSizeFilterSurf.setInputCloud(m_feats_undistort); m_downSizeFilterSurf.filter(*m_feats_down_body);
// (/root/reference/src/voxel_mapping.cpp:1715, :1888-1889; mesher copy: src/ImMesh_mesh_reconstruction.cpp:335-338).
// Semantics = published algorithm of pcl::VoxelGrid<PointT>::applyFilter with default parameters (restated in
// oracle/orc_frontend.hpp, which this path matches bit for bit): float inverse leaf, float min/max box, int cell index
// idx = ijk . (1, dx, dx*dy), points grouped by idx in ascending order, float centroid per group with the points summed in
// ascending scan order (PCL's std::sort leaves that order unspecified; both sides define it as the stable order).
//
// Kernels (all HBM-streaming, 12 B in per point per pass):
//   k_vg_minmax   block min/max of the finite points -> ordered-int atomics               12 n bytes in
//   k_vg_setup    one thread: box -> min_b, div_b, multipliers, PCL's INT_MAX overflow test
//   k_vg_keys     idx per point (0xFFFFFFFF for non-finite points, sorted last and dropped)  12 n in, 8 n out
//   k_rs_hist / k_rs_scan / k_rs_scatter  x4   stable LSD radix sort of (idx, point index), 8-bit digits, 2048-element tiles;
//                 ranks inside a tile from __match_any_sync (warp) + per-warp digit counts (shared memory)   16 n in+out per pass
//   k_vg_heads / k_rs_scan / k_vg_centroid   run heads -> output slots (tile counts + scan), serial float sums per run
// There is no CPU path: the entry points need a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/immesh_b200.h"
#include "common_host.hpp"
#include "handles.hpp"

using immesh::im_fail;

namespace {

#include "radix_sort.cuh"


struct VgGrid {
    int mn[3], mx[3];        // ordered-int encodings of the float box (atomicMin / atomicMax targets)
    int min_b[3], div_b[3];
    int mul[3];
    int too_small;           // PCL: "Leaf size is too small for the input dataset" -> output = input
    int n_finite;
    int m_out;
};

__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ bool finite3(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

__global__ void k_vg_reset(VgGrid* g) {
    if (threadIdx.x < 3) { g->mn[threadIdx.x] = 0x7fffffff; g->mx[threadIdx.x] = (int)0x80000000; }
    if (threadIdx.x == 0) { g->too_small = 0; g->n_finite = 0; g->m_out = 0; }
}
__global__ void __launch_bounds__(VG_THREADS) k_vg_minmax(const float* __restrict__ pts, int n, VgGrid* g) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        if (!finite3(x, y, z)) continue;
        mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
        ++cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(&g->mn[a], f2ord(mn[a])); atomicMax(&g->mx[a], f2ord(mx[a])); }
        atomicAdd(&g->n_finite, cnt);
    }
}
// applyFilter: box -> grid (float arithmetic exactly as PCL: products with the float inverse leaf, floor, int casts)
__global__ void k_vg_setup(VgGrid* g, float inv) {
    if (threadIdx.x != 0) return;
    if (g->n_finite == 0) { g->div_b[0] = g->div_b[1] = g->div_b[2] = 0; return; }
    float mn[3], mx[3];
    for (int a = 0; a < 3; ++a) { mn[a] = ord2f(g->mn[a]); mx[a] = ord2f(g->mx[a]); }
    const long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1, dz = (long long)((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > 2147483647LL) { g->too_small = 1; return; }
    for (int a = 0; a < 3; ++a) {
        g->min_b[a] = (int)floorf(mn[a] * inv);
        const int max_b = (int)floorf(mx[a] * inv);
        g->div_b[a] = max_b - g->min_b[a] + 1;
    }
    g->mul[0] = 1; g->mul[1] = g->div_b[0]; g->mul[2] = g->div_b[0] * g->div_b[1];
}
__global__ void __launch_bounds__(VG_THREADS) k_vg_keys(const float* __restrict__ pts, int n, const VgGrid* g, float inv, unsigned int* keys, unsigned int* vals) {
    const int mb0 = g->min_b[0], mb1 = g->min_b[1], mb2 = g->min_b[2], m1 = g->mul[1], m2 = g->mul[2];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        unsigned int key = 0xFFFFFFFFu;
        if (finite3(x, y, z)) {
            const int i0 = (int)(floorf(x * inv) - (float)mb0), i1 = (int)(floorf(y * inv) - (float)mb1), i2 = (int)(floorf(z * inv) - (float)mb2);
            key = (unsigned int)(i0 + i1 * m1 + i2 * m2);
        }
        keys[i] = key;
        vals[i] = (unsigned int)i;
    }
}

// ---- runs of equal idx -> centroids
__global__ void __launch_bounds__(VG_THREADS) k_vg_heads(const unsigned int* __restrict__ keys, int n, int* tile_heads) {
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int base = blockIdx.x * VG_TILE;
    int c = 0;
    for (int r = 0; r < VG_TILE / VG_THREADS; ++r) {
        const int i = base + r * VG_THREADS + threadIdx.x;
        if (i < n) {
            const unsigned int k = keys[i];
            if (k != 0xFFFFFFFFu && (i == 0 || keys[i - 1] != k)) ++c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_heads[blockIdx.x] = s_cnt;
}
// head_pos != nullptr: only record where run `slot` starts (the sums are then taken by k_vg_centroid_warp)
__global__ void __launch_bounds__(VG_THREADS) k_vg_centroid(const float* __restrict__ pts, const unsigned int* __restrict__ keys, const unsigned int* __restrict__ vals, int n,
                                                             const int* __restrict__ tile_base, float* out, int* head_pos) {
    __shared__ int s_warp[VG_THREADS / 32];
    __shared__ int s_run;
    if (threadIdx.x == 0) s_run = tile_base[blockIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = blockIdx.x * VG_TILE;
    for (int r = 0; r < VG_TILE / VG_THREADS; ++r) {
        const int i = base + r * VG_THREADS + threadIdx.x;
        unsigned int k = 0xFFFFFFFFu;
        bool head = false;
        if (i < n) { k = keys[i]; head = k != 0xFFFFFFFFu && (i == 0 || keys[i - 1] != k); }
        const unsigned int hb = __ballot_sync(0xffffffffu, head);
        if (lane == 0) s_warp[warp] = __popc(hb);
        __syncthreads();
        int slot = s_run + __popc(hb & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) slot += s_warp[w];
        if (head && head_pos) {
            head_pos[slot] = i;
        } else if (head) {   // CentroidPoint: float sums in run order, divided by (float)count
            float sx = 0.f, sy = 0.f, sz = 0.f;
            int e = i;
            while (e < n && keys[e] == k) {
                const unsigned int p = vals[e];
                sx += pts[3 * (size_t)p]; sy += pts[3 * (size_t)p + 1]; sz += pts[3 * (size_t)p + 2];
                ++e;
            }
            const float cnt = (float)(e - i);
            out[3 * (size_t)slot] = sx / cnt; out[3 * (size_t)slot + 1] = sy / cnt; out[3 * (size_t)slot + 2] = sz / cnt;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < VG_THREADS / 32; ++w) t += s_warp[w]; s_run += t; }
        __syncthreads();
    }
}
// one warp per run: the lanes fetch 32 points of the run at once (the loads are what made the thread-per-run loop slow: a dense
// leaf near the sensor holds hundreds of points, ~1 us of dependent misses each), then every lane adds the 32 values in
// run order from shuffles -- the float sum keeps the serial order of CentroidPoint, only the memory latency is taken off the chain.
__global__ void __launch_bounds__(VG_THREADS) k_vg_centroid_warp(const float* __restrict__ pts, const unsigned int* __restrict__ vals, const int* __restrict__ head_pos,
                                                                  const int* __restrict__ total_heads, const VgGrid* g, float* out) {
    const int lane = threadIdx.x & 31;
    const int m = *total_heads, n_valid = g->n_finite;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); slot < m; slot += nwarps) {
        const int b = head_pos[slot], e = (slot + 1 < m) ? head_pos[slot + 1] : n_valid;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int c = b; c < e; c += 32) {
            const int i = c + lane;
            float x = 0.f, y = 0.f, z = 0.f;
            if (i < e) { const unsigned int p = vals[i]; x = pts[3 * (size_t)p]; y = pts[3 * (size_t)p + 1]; z = pts[3 * (size_t)p + 2]; }
            const int cnt = min(32, e - c);
            for (int t = 0; t < cnt; ++t) {
                sx += __shfl_sync(0xffffffffu, x, t); sy += __shfl_sync(0xffffffffu, y, t); sz += __shfl_sync(0xffffffffu, z, t);
            }
        }
        if (lane == 0) {
            const float cnt = (float)(e - b);
            out[3 * (size_t)slot] = sx / cnt; out[3 * (size_t)slot + 1] = sy / cnt; out[3 * (size_t)slot + 2] = sz / cnt;
        }
    }
}
__global__ void k_vg_finish(VgGrid* g, const int* total_heads, int n) {
    if (threadIdx.x == 0) g->m_out = g->too_small ? n : *total_heads;
}
__global__ void __launch_bounds__(VG_THREADS) k_vg_copy_if_small(const float* __restrict__ pts, int n, const VgGrid* g, float* out) {
    if (!g->too_small) return;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 3 * (size_t)n; i += (size_t)gridDim.x * blockDim.x) out[i] = pts[i];
}

// front of the front-end: strided input (3 floats xyz, or 4 as the IMU stage emits: xyz + curvature) -> packed [n][3], with the
// KITTI laser calibration (voxel_mapping.cpp:1844-1859) applied on the way when preprocess/calib_laser is set
__global__ void __launch_bounds__(VG_THREADS) k_frontend_pack(const float* __restrict__ src, int n, int stride, int calib, float* __restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float x = src[(size_t)i * stride], y = src[(size_t)i * stride + 1], z = src[(size_t)i * stride + 2];
        if (calib) immesh::kitti_calib_point(&x, &y, &z);
        dst[3 * (size_t)i] = x; dst[3 * (size_t)i + 1] = y; dst[3 * (size_t)i + 2] = z;
    }
}

}  // namespace

struct immesh_voxelgrid {
    int max_points = 0, nblocks_max = 0, n_sm = 148;
    cudaStream_t stream = nullptr;
    float* d_in = nullptr;       // staging for host input / packed (calibrated) cloud of immesh_frontend_prepare: the current slot of
    float* d_in_ring[8] = {};   // a ring of 8 (a mesh frame may still be reading the cloud of up to IM_SLOTS + 1 scans ago)
    int in_idx = 0;
    float* d_raw = nullptr;      // [max_points][4] strided input of immesh_frontend_prepare
    float* h_raw = nullptr;      // pinned staging for it
    cudaEvent_t ev_raw = nullptr; // the H2D copy out of h_raw has completed
    int raw_pending = 0;
    float* d_out = nullptr;      // [max_points][3]
    unsigned int *d_k[2] = {nullptr, nullptr}, *d_v[2] = {nullptr, nullptr};
    int* d_hist = nullptr;       // [256 * nblocks_max]
    int* d_tile = nullptr;       // [nblocks_max + 1]
    int* d_head = nullptr;       // [max_points] start of every run in the sorted order
    int warp_centroid = 1;       // 1: warp-per-run sums (k_vg_centroid_warp); 0: thread-per-run loop (IMMESH_VG_WARP_CENTROID=0)
    VgGrid* d_grid = nullptr;
    VgGrid* h_grid = nullptr;    // pinned
    float* h_pts = nullptr;      // pinned staging [max_points][3]
    int last_m = 0;
};

extern "C" {

int immesh_voxelgrid_create(int max_points, immesh_voxelgrid_t** out) {
    if (!out || max_points < 1) return im_fail(IMMESH_E_INVALID, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return im_fail(IMMESH_E_NO_DEVICE, "no CUDA device: immesh_b200 has no CPU path"); }
    immesh_voxelgrid* h = new immesh_voxelgrid();
    h->max_points = max_points;
    h->nblocks_max = (max_points + VG_TILE - 1) / VG_TILE;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, dev);
    IM_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const size_t n = (size_t)max_points;
    for (int i = 0; i < 8; ++i) IM_CUDA(cudaMalloc((void**)&h->d_in_ring[i], n * 3 * sizeof(float)));
    h->d_in = h->d_in_ring[0];
    IM_CUDA(cudaMalloc((void**)&h->d_out, n * 3 * sizeof(float)));
    IM_CUDA(cudaMalloc((void**)&h->d_raw, n * 4 * sizeof(float)));
    IM_CUDA(cudaMallocHost((void**)&h->h_raw, n * 4 * sizeof(float)));
    IM_CUDA(cudaEventCreateWithFlags(&h->ev_raw, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        IM_CUDA(cudaMalloc((void**)&h->d_k[i], n * sizeof(unsigned int)));
        IM_CUDA(cudaMalloc((void**)&h->d_v[i], n * sizeof(unsigned int)));
    }
    IM_CUDA(cudaMalloc((void**)&h->d_hist, (size_t)256 * h->nblocks_max * sizeof(int)));
    IM_CUDA(cudaMalloc((void**)&h->d_tile, ((size_t)h->nblocks_max + 1) * sizeof(int)));
    IM_CUDA(cudaMalloc((void**)&h->d_head, n * sizeof(int)));
    if (const char* e = std::getenv("IMMESH_VG_WARP_CENTROID")) h->warp_centroid = (e[0] == '1') ? 1 : 0;
    IM_CUDA(cudaMalloc((void**)&h->d_grid, sizeof(VgGrid)));
    IM_CUDA(cudaMallocHost((void**)&h->h_grid, sizeof(VgGrid)));
    IM_CUDA(cudaMallocHost((void**)&h->h_pts, n * 3 * sizeof(float)));
    *out = h;
    return IMMESH_OK;
}

int immesh_voxelgrid_destroy(immesh_voxelgrid_t* h) {
    if (!h) return IMMESH_OK;
    cudaStreamSynchronize(h->stream);
    for (int i = 0; i < 8; ++i) cudaFree(h->d_in_ring[i]);
    cudaFree(h->d_out); cudaFree(h->d_raw); cudaFreeHost(h->h_raw);
    if (h->ev_raw) cudaEventDestroy(h->ev_raw);
    for (int i = 0; i < 2; ++i) { cudaFree(h->d_k[i]); cudaFree(h->d_v[i]); }
    cudaFree(h->d_hist); cudaFree(h->d_tile); cudaFree(h->d_head); cudaFree(h->d_grid);
    cudaFreeHost(h->h_grid); cudaFreeHost(h->h_pts);
    cudaStreamDestroy(h->stream);
    delete h;
    return IMMESH_OK;
}

// queue the filter's launch sequence for the packed device cloud d_pts on stream st (no host synchronisation)
static int vg_enqueue(immesh_voxelgrid* h, const float* d_pts, int n, float leaf, cudaStream_t st) {
    const float inv = 1.0f / leaf;
    const int nb = (n + VG_TILE - 1) / VG_TILE;
    const int gs = std::min(nb * (VG_TILE / VG_THREADS), h->n_sm * 8);
    IM_LAUNCH(k_vg_reset, 1, 32, 0, st, h->d_grid);
    IM_LAUNCH(k_vg_minmax, gs, VG_THREADS, 0, st, d_pts, n, h->d_grid);
    IM_LAUNCH(k_vg_setup, 1, 32, 0, st, h->d_grid, inv);
    IM_LAUNCH(k_vg_keys, gs, VG_THREADS, 0, st, d_pts, n, (const VgGrid*)h->d_grid, inv, h->d_k[0], h->d_v[0]);
    int cur = 0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        IM_LAUNCH(k_rs_hist, nb, VG_THREADS, 0, st, (const unsigned int*)h->d_k[cur], n, shift, h->d_hist, nb);
        IM_LAUNCH(k_rs_scan, 1, 1024, 0, st, h->d_hist, 256 * nb, (int*)nullptr);
        IM_LAUNCH(k_rs_scatter, nb, VG_THREADS, 0, st, (const unsigned int*)h->d_k[cur], (const unsigned int*)h->d_v[cur], h->d_k[cur ^ 1], h->d_v[cur ^ 1], n, shift,
                  (const int*)h->d_hist, nb);
        cur ^= 1;
    }
    IM_LAUNCH(k_vg_heads, nb, VG_THREADS, 0, st, (const unsigned int*)h->d_k[cur], n, h->d_tile);
    IM_LAUNCH(k_rs_scan, 1, 1024, 0, st, h->d_tile, nb, h->d_tile + h->nblocks_max);
    if (h->warp_centroid) {
        IM_LAUNCH(k_vg_centroid, nb, VG_THREADS, 0, st, d_pts, (const unsigned int*)h->d_k[cur], (const unsigned int*)h->d_v[cur], n, (const int*)h->d_tile, h->d_out, h->d_head);
        IM_LAUNCH(k_vg_centroid_warp, h->n_sm * 4, VG_THREADS, 0, st, d_pts, (const unsigned int*)h->d_v[cur], (const int*)h->d_head,
                  (const int*)(h->d_tile + h->nblocks_max), (const VgGrid*)h->d_grid, h->d_out);
    } else {
        IM_LAUNCH(k_vg_centroid, nb, VG_THREADS, 0, st, d_pts, (const unsigned int*)h->d_k[cur], (const unsigned int*)h->d_v[cur], n, (const int*)h->d_tile, h->d_out, (int*)nullptr);
    }
    IM_LAUNCH(k_vg_finish, 1, 32, 0, st, h->d_grid, (const int*)(h->d_tile + h->nblocks_max), n);
    IM_LAUNCH(k_vg_copy_if_small, gs, VG_THREADS, 0, st, d_pts, n, (const VgGrid*)h->d_grid, h->d_out);
    IM_CUDA(cudaGetLastError());
    return IMMESH_OK;
}

int immesh_voxelgrid_filter(immesh_voxelgrid_t* h, const float* xyz, int n, int on_device, float leaf, float* out_xyz, int* m_out, int* leaf_too_small) {
    if (!h || (!xyz && n > 0) || n < 0 || !(leaf > 0.f)) return im_fail(IMMESH_E_INVALID, "bad argument");
    if (n > h->max_points) return im_fail(IMMESH_E_CAPACITY, "cloud larger than max_points");
    if (m_out) *m_out = 0;
    if (leaf_too_small) *leaf_too_small = 0;
    h->last_m = 0;
    if (n == 0) return IMMESH_OK;
    cudaStream_t st = h->stream;
    const float* d_pts = xyz;
    if (!on_device) {
        std::memcpy(h->h_pts, xyz, (size_t)n * 3 * sizeof(float));
        IM_CUDA(cudaMemcpyAsync(h->d_in, h->h_pts, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
        d_pts = h->d_in;
    }
    const int rc = vg_enqueue(h, d_pts, n, leaf, st);
    if (rc) return rc;
    IM_CUDA(cudaMemcpyAsync(h->h_grid, h->d_grid, sizeof(VgGrid), cudaMemcpyDeviceToHost, st));
    IM_CUDA(cudaStreamSynchronize(st));
    const int m = h->h_grid->m_out;
    h->last_m = m;
    if (m_out) *m_out = m;
    if (leaf_too_small) *leaf_too_small = h->h_grid->too_small;
    if (out_xyz && m > 0) {
        IM_CUDA(cudaMemcpyAsync(h->h_pts, h->d_out, (size_t)m * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
        IM_CUDA(cudaStreamSynchronize(st));
        std::memcpy(out_xyz, h->h_pts, (size_t)m * 3 * sizeof(float));
    }
    return IMMESH_OK;
}

// upload (host input) + pack / calibrate into d_in on stream st
static int frontend_pack(immesh_voxelgrid* h, const float* pts, int n, int stride, int on_device, int calib_laser, cudaStream_t st) {
    h->in_idx = (h->in_idx + 1) & 7;
    h->d_in = h->d_in_ring[h->in_idx];
    const float* d_src = pts;
    if (!on_device) {
        cudaPointerAttributes a;
        const float* src = pts;
        if (cudaPointerGetAttributes(&a, pts) != cudaSuccess || a.type != cudaMemoryTypeHost) {
            cudaGetLastError();
            if (h->raw_pending) { IM_CUDA(cudaEventSynchronize(h->ev_raw)); h->raw_pending = 0; }   // the previous copy out of the staging buffer
            std::memcpy(h->h_raw, pts, (size_t)n * stride * sizeof(float));
            src = h->h_raw;
        }
        IM_CUDA(cudaMemcpyAsync(h->d_raw, src, (size_t)n * stride * sizeof(float), cudaMemcpyHostToDevice, st));
        if (src == h->h_raw) { IM_CUDA(cudaEventRecord(h->ev_raw, st)); h->raw_pending = 1; }
        d_src = h->d_raw;
    }
    IM_LAUNCH(k_frontend_pack, std::min((n + VG_THREADS - 1) / VG_THREADS, h->n_sm * 8), VG_THREADS, 0, st, d_src, n, stride, calib_laser, h->d_in);
    return IMMESH_OK;
}

int immesh_frontend_prepare(immesh_voxelgrid_t* h, const float* pts, int n, int stride, int on_device, int calib_laser, float* out_xyz) {
    if (!h || (!pts && n > 0) || n < 0 || (stride != 3 && stride != 4)) return im_fail(IMMESH_E_INVALID, "bad argument (stride must be 3 or 4)");
    if (n > h->max_points) return im_fail(IMMESH_E_CAPACITY, "cloud larger than max_points");
    if (n == 0) return IMMESH_OK;
    int rc = frontend_pack(h, pts, n, stride, on_device, calib_laser, h->stream);
    if (rc) return rc;
    IM_CUDA(cudaGetLastError());
    if (out_xyz) {
        IM_CUDA(cudaMemcpyAsync(h->h_pts, h->d_in, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        IM_CUDA(cudaStreamSynchronize(h->stream));
        std::memcpy(out_xyz, h->h_pts, (size_t)n * 3 * sizeof(float));
    } else {
        IM_CUDA(cudaStreamSynchronize(h->stream));
    }
    return IMMESH_OK;
}
const float* immesh_voxelgrid_input_points(immesh_voxelgrid_t* h) { return h ? h->d_in : nullptr; }

// The device-resident front-end chain of one scan, queued on the localization handle's stream with no host round
// trip: [KITTI laser calibration] -> pcl::VoxelGrid -> predict + IESKF iterations + map update.  The down-sampled cloud and its
// size never leave the device (the step reads the count from the filter's result block).
int immesh_lio_step_async_raw(immesh_lio_t* lio, immesh_voxelgrid_t* h, const float* pts, int n, int stride, int on_device, int calib_laser, float leaf,
                              double dt, double cov_gyr, double cov_acc) {
    if (!lio || !h || (!pts && n > 0) || n < 1 || (stride != 3 && stride != 4) || !(leaf > 0.f)) return im_fail(IMMESH_E_INVALID, "bad argument");
    if (n > h->max_points) return im_fail(IMMESH_E_CAPACITY, "cloud larger than max_points");
    cudaStream_t st = lio->stream;
    int rc = frontend_pack(h, pts, n, stride, on_device, calib_laser, st);
    if (rc) return rc;
    rc = vg_enqueue(h, h->d_in, n, leaf, st);
    if (rc) return rc;
    return immesh_lio_step_async_dev_n(lio, h->d_out, n < lio->max_scan ? n : lio->max_scan, &h->d_grid->m_out, dt, cov_gyr, cov_acc);
}

const float* immesh_voxelgrid_device_points(immesh_voxelgrid_t* h) { return h ? h->d_out : nullptr; }

}  // extern "C"
