// immesh_b200 -- NCCL, bound at run time (dlopen) so that the single-GPU library has no link-time dependency on it.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstddef>

namespace immesh {
typedef struct { char internal[128]; } nccl_uid_t;
typedef void* nccl_comm_t;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_uid_t*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid_t, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        GetUniqueId = (int (*)(nccl_uid_t*))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (int (*)(nccl_comm_t*, int, nccl_uid_t, int))dlsym(lib, "ncclCommInitRank");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(lib, "ncclAllReduce");
        AllGather = (int (*)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t))dlsym(lib, "ncclAllGather");
        CommDestroy = (int (*)(nccl_comm_t))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && AllReduce && AllGather && CommDestroy;
    }
};
inline NcclApi& nccl() { static NcclApi a; return a; }
const int kNcclUint8 = 1, kNcclUint32 = 3, kNcclUint64 = 5, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values (nccl.h)
}  // namespace immesh
