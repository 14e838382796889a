// immesh_b200 -- peer windows: one cudaMalloc'ed buffer per rank, mapped into every other rank of the node through CUDA IPC
// (NVLink / NVSwitch peer memory).  The sharded paths exchange their per-iteration / per-frame data by writing it straight
// into the consumer ranks' windows from the kernel that produced it and raising a per-source epoch flag (release, system
// scope); the consuming kernel spins on the flags of its peers (acquire, system scope).  No collective call, no host
// involvement, no extra launch: the transfer is part of the producing kernel's tail and the consuming kernel's head.
//
// Ordering protocol (every use below follows it):
//   producer block : remote stores -> __threadfence_system() -> __syncthreads() -> block-done counter (device scope)
//   last block     : __threadfence_system() -> st.release.sys flag[src] = epoch   (one flag per consumer rank)
//   consumer       : ld.acquire.sys flag[src] >= epoch for every src -> __syncthreads() -> reads bypass L1 (__ldcg)
// Epochs are a host-side counter advanced identically on all ranks (they run the same launch sequence), flags only grow,
// nothing is ever reset, so a rank that is ahead can never be confused with one that is behind.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

#include "nccl_api.hpp"

#define IM_MAX_RANKS 8

namespace immesh {

#if defined(__CUDACC__)
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// spin until *flag >= epoch.  Bounded (default ~10 s of SM clocks, IMMESH_PEER_TIMEOUT_MS): a dead peer must not hang the
// GPU.  On expiry `bit` is OR-ed into *err (the handle's error word, reported by the next wait/step call); once it is set,
// later waits return at once, so a lost peer costs ONE timeout, not one per kernel.
static __device__ long long g_peer_timeout_cycles = 20000000000LL;
__device__ __forceinline__ bool wait_epoch(const unsigned long long* flag, unsigned long long epoch, int* err, int bit) {
    if (*(volatile int*)err & bit) return false;
    const long long t0 = clock64(), limit = g_peer_timeout_cycles;
    while (ld_acquire_sys(flag) < epoch) {
        __nanosleep(40);
        if (clock64() - t0 > limit) { atomicOr(err, bit); return false; }
    }
    return true;
}
#endif

struct PeerWindow {
    unsigned char* local = nullptr;
    unsigned char* peer[IM_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // peer[rank] == local
    size_t bytes = 0;
    int rank = 0, n = 1;
    bool ok = false;
    unsigned long long epoch = 0;   // host counter, advanced once per exchange
};

// Allocates the local window, exchanges the IPC handles through the (already initialised) NCCL communicator and maps the
// peers' windows.  Returns cudaSuccess or the failing CUDA error; `w.ok` is only set when every peer is mapped.
inline cudaError_t peer_window_open(PeerWindow& w, size_t bytes, int rank, int n, nccl_comm_t comm, cudaStream_t st) {
    w.ok = false;
    w.rank = rank;
    w.n = n;
    w.bytes = bytes;
    if (n > IM_MAX_RANKS) return cudaErrorInvalidValue;
    cudaError_t e = cudaMalloc((void**)&w.local, bytes);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemset(w.local, 0, bytes)) != cudaSuccess) return e;
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;
    cudaIpcMemHandle_t mine;
    if ((e = cudaIpcGetMemHandle(&mine, w.local)) != cudaSuccess) return e;
    unsigned char* d_x = nullptr;
    if ((e = cudaMalloc((void**)&d_x, sizeof(mine) * (size_t)(n + 1))) != cudaSuccess) return e;
    cudaMemcpy(d_x, &mine, sizeof(mine), cudaMemcpyHostToDevice);
    if (nccl().AllGather(d_x, d_x + sizeof(mine), sizeof(mine), kNcclUint8, comm, st)) { cudaFree(d_x); return cudaErrorUnknown; }
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { cudaFree(d_x); return e; }
    cudaIpcMemHandle_t all[IM_MAX_RANKS];
    cudaMemcpy(all, d_x + sizeof(mine), sizeof(mine) * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(d_x);
    for (int r = 0; r < n; ++r) {
        if (r == rank) { w.peer[r] = w.local; continue; }
        void* p = nullptr;
        if ((e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) return e;
        w.peer[r] = (unsigned char*)p;
    }
#if defined(__CUDACC__)
    if (const char* t = std::getenv("IMMESH_PEER_TIMEOUT_MS")) {
        const long long cyc = (long long)(std::atof(t) * 2.0e6);   // ~2 GHz SM clock
        if (cyc > 0) cudaMemcpyToSymbol(g_peer_timeout_cycles, &cyc, sizeof(cyc));
    }
#endif
    w.ok = true;
    return cudaSuccess;
}
inline void peer_window_close(PeerWindow& w) {
    for (int r = 0; r < w.n; ++r)
        if (r != w.rank && w.peer[r]) cudaIpcCloseMemHandle(w.peer[r]);
    if (w.local) cudaFree(w.local);
    w = PeerWindow();
}

}  // namespace immesh
