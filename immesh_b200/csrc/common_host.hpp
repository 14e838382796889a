// immesh_b200 -- host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <string>

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
}  // namespace immesh

#define IM_CUDA(expr)                                                              \
    do {                                                                           \
        cudaError_t im_e_ = (expr);                                                \
        if (im_e_ != cudaSuccess) return immesh::im_fail_cuda(im_e_, __FILE__, __LINE__); \
    } while (0)
