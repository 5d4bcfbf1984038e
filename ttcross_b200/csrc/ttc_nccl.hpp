// =============================================================================
// ttc_nccl.hpp — the handful of NCCL entry points the sweep's neighbour exchange uses, bound at run time.
//
// The library must load (and run single-GPU problems) in a plain Fortran/C process with no NCCL installed, and must
// share ONE NCCL instance with a host process that already has one (e.g. torch.distributed): so libnccl.so.2 is
// dlopen'ed on first use — an already-loaded copy with that soname is reused — instead of being a link-time dependency.
// Prototypes restated from the public NCCL 2.x API (nccl.h); only plain C types cross the boundary.
// =============================================================================
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string>

namespace ttc {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
enum { NCCL_INT8 = 0, NCCL_FLOAT64 = 8 };     // ncclDataType_t values of nccl.h (ncclInt8 = 0, ncclFloat64 = 8)

struct NcclApi {
    void* so = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    std::string err;

    bool load() {
        if (so) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (so) break; }
        if (!so) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
        auto sym = [&](const char* nm) { void* p = dlsym(so, nm); if (!p && err.empty()) err = std::string("libnccl lacks ") + nm; return p; };
        GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
        CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))sym("ncclCommInitRank");
        CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
        AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))sym("ncclAllGather");
        Send = (int (*)(const void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclSend");
        Recv = (int (*)(void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclRecv");
        GroupStart = (int (*)())sym("ncclGroupStart");
        GroupEnd = (int (*)())sym("ncclGroupEnd");
        GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
        GetVersion = (int (*)(int*))sym("ncclGetVersion");
        if (!err.empty()) { dlclose(so); so = nullptr; return false; }
        return true;
    }
};
inline NcclApi& nccl_api() { static NcclApi a; return a; }

}  // namespace ttc
