// immesh_b200 -- host-side helpers shared by the C-ABI translation units: error reporting, launch accounting
// and an optional CUDA-event profiler (per-kernel device time on the launching stream, used by bench.py for the
// roofline numbers; off by default so that timed runs carry no event overhead).
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <map>
#include <string>
#include <tuple>
#include <cstdlib>
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
    // optional timeline: start/end of every launch relative to a base event (diagnosis of the two-stream pipeline)
    bool timeline_on = false;
    cudaEvent_t base = nullptr;
    struct Span { const char* name; float t0, t1; };
    std::vector<Span> timeline;
    void start_timeline() {
        if (!base) cudaEventCreate(&base);
        cudaDeviceSynchronize();
        cudaEventRecord(base, 0);
        cudaEventSynchronize(base);
        timeline.clear();
        timeline_on = true;
    }
    void collect() {  // harvests every launch whose end event has completed; the others stay pending
        std::vector<Rec> keep;
        for (Rec& r : pending) {
            if (cudaEventQuery(r.b) == cudaErrorNotReady) { keep.push_back(r); continue; }
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
                auto& t = totals[r.name];
                t.first += ms;
                t.second += 1;
                if (timeline_on) {
                    Span sp{r.name, 0.f, 0.f};
                    cudaEventElapsedTime(&sp.t0, base, r.a);
                    cudaEventElapsedTime(&sp.t1, base, r.b);
                    timeline.push_back(sp);
                }
            }
            pool.push_back(r.a);
            pool.push_back(r.b);
        }
        pending.swap(keep);
        cudaGetLastError();
    }
};
Profiler& profiler();  // defined in misc_capi.cu
}  // namespace immesh

#define IM_CUDA(expr)                                                              \
    do {                                                                           \
        cudaError_t im_e_ = (expr);                                                \
        if (im_e_ != cudaSuccess) return immesh::im_fail_cuda(im_e_, __FILE__, __LINE__); \
    } while (0)

// ---- CUDA-graph replay of a fixed launch sequence.  The per-scan work is ~40 short kernels, so queueing them one by
// one is host-launch-bound.  The same host code path runs in three modes: direct launch; capture (first pipelined call:
// the launches are recorded into a graph and the node of every launch is remembered); update (later calls: every
// "launch" only rewrites the parameters / grid of its node in the instantiated graph, then ONE cudaGraphLaunch replays it).
namespace immesh {
struct GraphCtx {
    int mode = 0;                      // 0 direct, 1 capturing, 2 updating node parameters
    cudaGraphExec_t exec = nullptr;
    cudaGraph_t graph = nullptr;
    std::vector<cudaGraphNode_t> nodes;
    size_t cursor = 0;
    unsigned sig = 0;                  // control-flow signature the graph was captured with
    cudaError_t err = cudaSuccess;
    long long captures = 0, replays = 0, failures = 0;
    void destroy() {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        exec = nullptr; graph = nullptr; nodes.clear();
    }
};
inline GraphCtx*& graph_ctx() {
    static thread_local GraphCtx* g = nullptr;
    return g;
}
template <typename... KArgs, typename... Args>
inline void im_launch(const char* name, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    Profiler& pr = profiler();
    pr.launches++;
    std::tuple<KArgs...> vals(std::forward<Args>(args)...);   // exactly the kernel's parameter types
    void* ptrs[sizeof...(KArgs) ? sizeof...(KArgs) : 1];
    {
        size_t k = 0;
        std::apply([&](auto&... v) { ((ptrs[k++] = (void*)&v), ...); }, vals);
    }
    GraphCtx* g = graph_ctx();
    if (g && g->mode == 2) {
        cudaKernelNodeParams kp{};
        kp.func = (void*)kern; kp.gridDim = grid; kp.blockDim = block; kp.sharedMemBytes = (unsigned)smem;
        kp.kernelParams = ptrs; kp.extra = nullptr;
        cudaError_t e = (g->cursor < g->nodes.size()) ? cudaGraphExecKernelNodeSetParams(g->exec, g->nodes[g->cursor], &kp) : cudaErrorInvalidValue;
        g->cursor++;
        if (e != cudaSuccess && g->err == cudaSuccess) g->err = e;
        return;
    }
    if (pr.enabled) pr.begin(name, st);
    cudaError_t e = cudaLaunchKernel((const void*)kern, grid, block, ptrs, smem, st);
    if (pr.enabled) pr.end(st);
    if (g && g->mode == 1) {
        cudaStreamCaptureStatus status;
        const cudaGraphNode_t* deps = nullptr;
        size_t ndeps = 0;
        cudaError_t e2 = cudaStreamGetCaptureInfo(st, &status, nullptr, nullptr, &deps, &ndeps);
        if (e == cudaSuccess && e2 == cudaSuccess && status == cudaStreamCaptureStatusActive && ndeps == 1) g->nodes.push_back(deps[0]);
        else if (g->err == cudaSuccess) g->err = (e != cudaSuccess) ? e : (e2 != cudaSuccess ? e2 : cudaErrorUnknown);
    }
}
inline bool im_replaying() { GraphCtx* g = graph_ctx(); return g && g->mode == 2; }
// queue `body` (a fixed sequence of IM_LAUNCH + fork/join event calls guarded by !im_replaying()) on `st` through graph `g`.
// Returns cudaSuccess when the work was queued by a graph launch; on any failure the graph is dropped and the caller
// queues the work directly.
template <typename Body>
inline cudaError_t run_graphed(GraphCtx& g, unsigned sig, cudaStream_t st, Body&& body) {
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (g.exec && g.sig != sig) g.destroy();
        g.err = cudaSuccess;
        graph_ctx() = &g;
        if (!g.exec) {
            g.nodes.clear();
            cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            if (e != cudaSuccess) { graph_ctx() = nullptr; return e; }
            g.mode = 1;
            body();
            g.mode = 0;
            graph_ctx() = nullptr;
            e = cudaStreamEndCapture(st, &g.graph);
            if (e == cudaSuccess && g.err != cudaSuccess) e = g.err;
            if (e == cudaSuccess) e = cudaGraphInstantiate(&g.exec, g.graph, 0);
            if (e != cudaSuccess) { g.destroy(); cudaGetLastError(); g.failures++; return e; }
            g.sig = sig;
            g.captures++;
        } else {
            g.mode = 2;
            g.cursor = 0;
            body();
            g.mode = 0;
            graph_ctx() = nullptr;
            if (g.err != cudaSuccess || g.cursor != g.nodes.size()) { g.destroy(); cudaGetLastError(); g.failures++; continue; }  // re-capture
            g.replays++;
        }
        return cudaGraphLaunch(g.exec, st);
    }
    return cudaErrorUnknown;
}
// Static form: the body's launches take every per-call value from device memory, so once captured the graph is replayed by a
// bare cudaGraphLaunch -- the body is not run again and no node is touched.  `sig` must cover everything that changes the
// sequence or its arguments (handle pointers, modes, iteration counts).
template <typename Body>
inline cudaError_t run_graphed_static(GraphCtx& g, unsigned sig, cudaStream_t st, Body&& body) {
    if (g.exec && g.sig != sig) g.destroy();
    if (!g.exec) {
        g.err = cudaSuccess;
        g.nodes.clear();
        graph_ctx() = &g;
        cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) { graph_ctx() = nullptr; return e; }
        g.mode = 1;
        body();
        g.mode = 0;
        graph_ctx() = nullptr;
        e = cudaStreamEndCapture(st, &g.graph);
        if (e == cudaSuccess && g.err != cudaSuccess) e = g.err;
        if (e == cudaSuccess) e = cudaGraphInstantiate(&g.exec, g.graph, 0);
        if (e != cudaSuccess) { g.destroy(); cudaGetLastError(); g.failures++; return e; }
        g.sig = sig;
        g.captures++;
    } else {
        g.replays++;
        profiler().launches += (long long)g.nodes.size();   // launch accounting: every kernel node of the graph runs
    }
    return cudaGraphLaunch(g.exec, st);
}
}  // namespace immesh

// kernel launch with accounting; wrap template kernels in parentheses: IM_LAUNCH((k<256>), grid, block, smem, stream, args...)
#define IM_LAUNCH(KERN, GRID, BLOCK, SMEM, STREAM, ...) immesh::im_launch(#KERN, KERN, dim3(GRID), dim3(BLOCK), (SMEM), (STREAM), ##__VA_ARGS__)

// stream priority from the environment (experiments): 0 = default (lowest), 1 = highest
inline int im_stream_priority(const char* env) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    const char* e = getenv(env);
    return (e && atoi(e) == 1) ? hi : lo;
}
