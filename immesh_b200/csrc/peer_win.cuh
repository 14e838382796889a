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

inline void peer_window_close(PeerWindow& w);
// Allocates the local window, exchanges the IPC handles through the (already initialised) NCCL communicator and maps the
// peers' windows.  The decision "peer windows or NCCL" must be the SAME on every rank (a rank spinning on epoch flags while
// another sits in an all-reduce would hang both), so every rank always takes part in both collectives below, a local failure
// travels as a marker, and `w.ok` is set only when EVERY rank mapped EVERY peer; otherwise everything opened here is released
// again on all ranks and they all use the NCCL transport.  Returns cudaSuccess or the first local CUDA error
// (cudaErrorNotReady: this rank was fine but a peer was not).
inline cudaError_t peer_window_open(PeerWindow& w, size_t bytes, int rank, int n, nccl_comm_t comm, cudaStream_t st) {
    struct Msg { cudaIpcMemHandle_t h; int ok; int pad_[15]; };
    w.ok = false;
    w.rank = rank;
    w.n = n;
    w.bytes = bytes;
    if (n > IM_MAX_RANKS) return cudaErrorInvalidValue;          // same n on every rank: a uniform decision
    cudaError_t first = cudaSuccess;
    auto note = [&](cudaError_t e) { if (e != cudaSuccess && first == cudaSuccess) first = e; return e == cudaSuccess; };
    // exchange buffer first: without it this rank could not even report its failure (a 1 KB allocation; if it fails the
    // process is out of device memory and the caller's create/shard call fails as a whole)
    unsigned char* d_x = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_x, sizeof(Msg) * (size_t)(n + 1) + 16);
    if (e != cudaSuccess) return e;
    Msg mine;
    std::memset(&mine, 0, sizeof(mine));
    if (note(cudaMalloc((void**)&w.local, bytes)) && note(cudaMemset(w.local, 0, bytes)) && note(cudaDeviceSynchronize()) &&
        note(cudaIpcGetMemHandle(&mine.h, w.local)))
        mine.ok = 1;
    cudaMemcpy(d_x, &mine, sizeof(mine), cudaMemcpyHostToDevice);
    bool coll_ok = nccl().AllGather(d_x, d_x + sizeof(Msg), sizeof(Msg), kNcclUint8, comm, st) == 0;
    coll_ok = coll_ok && cudaStreamSynchronize(st) == cudaSuccess;
    Msg all[IM_MAX_RANKS];
    std::memset(all, 0, sizeof(all));
    if (coll_ok) cudaMemcpy(all, d_x + sizeof(Msg), sizeof(Msg) * (size_t)n, cudaMemcpyDeviceToHost);
    bool everyone = coll_ok;
    for (int r = 0; r < n && everyone; ++r) everyone = all[r].ok != 0;
    unsigned int failed = 0;
    if (everyone) {
        for (int r = 0; r < n; ++r) {
            if (r == rank) { w.peer[r] = w.local; continue; }
            void* p = nullptr;
            if (!note(cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess))) { failed = 1; break; }
            w.peer[r] = (unsigned char*)p;
        }
    } else {
        failed = 1;
    }
    // second round: did every rank map every peer?  (sum of the failure markers)
    if (coll_ok) {
        unsigned int* d_f = (unsigned int*)(d_x + sizeof(Msg) * (size_t)(n + 1));
        cudaMemcpy(d_f, &failed, sizeof(failed), cudaMemcpyHostToDevice);
        unsigned int total = 1;
        if (nccl().AllReduce(d_f, d_f, 1, kNcclUint32, kNcclSum, comm, st) == 0 && cudaStreamSynchronize(st) == cudaSuccess)
            cudaMemcpy(&total, d_f, sizeof(total), cudaMemcpyDeviceToHost);
        failed = total;
    }
    cudaFree(d_x);
    if (failed) {
        const int keep_rank = w.rank, keep_n = w.n;
        peer_window_close(w);
        w.rank = keep_rank; w.n = keep_n;
        cudaGetLastError();
        return first != cudaSuccess ? first : cudaErrorNotReady;
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
