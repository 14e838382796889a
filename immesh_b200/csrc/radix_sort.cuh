// immesh_b200 -- stable LSD radix sort of (u32 key, u32 value) pairs, 8-bit digits, 2048-element tiles; ranks inside a tile from
// __match_any_sync (warp) + per-warp digit counts in shared memory.  Used by the pcl::VoxelGrid front-end (cell index -> point) and by
// the IMU undistortion (time stamp -> point).  Include inside an anonymous namespace of the translation unit that launches it.
#pragma once
#define VG_TILE 2048
#define VG_THREADS 256

// ---- stable LSD radix sort, 8-bit digits.  hist is digit-major: hist[d * nblocks + b]
__global__ void __launch_bounds__(VG_THREADS) k_rs_hist(const unsigned int* __restrict__ keys, int n, int shift, int* hist, int nblocks) {
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * VG_TILE;
    for (int r = 0; r < VG_TILE / VG_THREADS; ++r) {
        const int i = base + r * VG_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}
// exclusive scan of `len` ints in place by ONE block (len <= a few 100 k); total -> *total_out if not null
__global__ void __launch_bounds__(1024) k_rs_scan(int* a, int len, int* total_out) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < len; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < len ? a[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int incl = x + (warp > 0 ? s_warp[warp - 1] : 0) + s_carry;
        if (i < len) a[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = s_carry;
}
__global__ void __launch_bounds__(VG_THREADS) k_rs_scatter(const unsigned int* __restrict__ kin, const unsigned int* __restrict__ vin, unsigned int* kout, unsigned int* vout,
                                                            int n, int shift, const int* __restrict__ hist, int nblocks) {
    __shared__ int run[256];                      // next free slot of every digit for this tile (global position)
    __shared__ int wcnt[VG_THREADS / 32][256];    // per-warp digit counts of the current round
    run[threadIdx.x] = hist[threadIdx.x * nblocks + blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = blockIdx.x * VG_TILE;
    for (int r = 0; r < VG_TILE / VG_THREADS; ++r) {
        for (int w = 0; w < VG_THREADS / 32; ++w) wcnt[w][threadIdx.x] = 0;
        __syncthreads();
        const int i = base + r * VG_THREADS + threadIdx.x;
        const bool valid = i < n;
        unsigned int key = 0, val = 0;
        if (valid) { key = kin[i]; val = vin[i]; }
        const unsigned int d = valid ? ((key >> shift) & 255u) : (256u + (unsigned)lane);   // invalid lanes: singleton groups
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) wcnt[warp][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            int off = run[d];
            for (int w = 0; w < warp; ++w) off += wcnt[w][d];
            kout[off + rank] = key;
            vout[off + rank] = val;
        }
        __syncthreads();
        int tot = 0;
        for (int w = 0; w < VG_THREADS / 32; ++w) tot += wcnt[w][threadIdx.x];
        run[threadIdx.x] += tot;
        __syncthreads();
    }
}

