// immesh_b200 -- host-side helpers shared by the C-ABI translation units: error reporting, launch accounting
// and an optional CUDA-event profiler (per-kernel device time on the launching stream, used by bench.py for the
// roofline numbers; off by default so that timed runs carry no event overhead).
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/immesh_b200.h"

namespace immesh {
inline std::string& last_error_storage() {
    static thread_local std::string s;
    return s;
}
inline int im_fail(int code, const char* msg) {
    last_error_storage() = msg ? msg : "";
    return code;
}
inline int im_fail_cuda(cudaError_t e, const char* file, int line) {
    char buf[512];
    std::snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e), file, line);
    last_error_storage() = buf;
    return IMMESH_E_CUDA;
}

struct Profiler {
    bool enabled = false;
    long long launches = 0;
    struct Rec { const char* name; cudaEvent_t a, b; };
    std::vector<Rec> pending;
    std::vector<cudaEvent_t> pool;
    std::map<std::string, std::pair<double, long long>> totals;  // name -> (ms, launches)
    cudaEvent_t get_event() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void begin(const char* name, cudaStream_t st) {
        Rec r{name, get_event(), get_event()};
        cudaEventRecord(r.a, st);
        pending.push_back(r);
    }
    void end(cudaStream_t st) { cudaEventRecord(pending.back().b, st); }
    void collect() {  // call after the stream has been synchronised
        for (Rec& r : pending) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
                auto& t = totals[r.name];
                t.first += ms;
                t.second += 1;
            }
            pool.push_back(r.a);
            pool.push_back(r.b);
        }
        pending.clear();
    }
};
Profiler& profiler();  // defined in misc_capi.cu
}  // namespace immesh

#define IM_CUDA(expr)                                                              \
    do {                                                                           \
        cudaError_t im_e_ = (expr);                                                \
        if (im_e_ != cudaSuccess) return immesh::im_fail_cuda(im_e_, __FILE__, __LINE__); \
    } while (0)

// kernel launch with accounting; wrap template kernels in parentheses: IM_LAUNCH((k<256>), grid, block, smem, stream, args...)
#define IM_LAUNCH(KERN, GRID, BLOCK, SMEM, STREAM, ...)                \
    do {                                                               \
        immesh::Profiler& im_p_ = immesh::profiler();                  \
        im_p_.launches++;                                              \
        if (im_p_.enabled) im_p_.begin(#KERN, (STREAM));               \
        KERN<<<(GRID), (BLOCK), (SMEM), (STREAM)>>>(__VA_ARGS__);      \
        if (im_p_.enabled) im_p_.end((STREAM));                        \
    } while (0)
