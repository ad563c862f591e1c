// nccl_dyn.h -- NCCL bound at run time with dlopen.
//
// The library must load on a CPU-only box (the symbol-export test) and inside a
// process that already carries torch's bundled NCCL, so libnccl is not a link
// dependency: the handful of entry points used for the slab halo exchange and
// the Jmin/Jmax all-reduce are resolved on first use.  dlopen("libnccl.so.2")
// returns the copy already mapped into the process when there is one.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

#include <functional>
#include <mutex>

namespace ncclx {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4,
               ncclUint64 = 5, ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;

struct Api {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

inline void load(Api& a)
{
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* n : names) {
        a.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.so) break;
    }
    if (!a.so) return;
#define NCCLX_SYM(field, sym) *(void**)(&a.field) = dlsym(a.so, sym)
    NCCLX_SYM(GetUniqueId, "ncclGetUniqueId");
    NCCLX_SYM(CommInitRank, "ncclCommInitRank");
    NCCLX_SYM(CommInitAll, "ncclCommInitAll");
    NCCLX_SYM(CommDestroy, "ncclCommDestroy");
    NCCLX_SYM(GroupStart, "ncclGroupStart");
    NCCLX_SYM(GroupEnd, "ncclGroupEnd");
    NCCLX_SYM(Send, "ncclSend");
    NCCLX_SYM(Recv, "ncclRecv");
    NCCLX_SYM(AllReduce, "ncclAllReduce");
    NCCLX_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCLX_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.GroupStart &&
           a.GroupEnd && a.Send && a.Recv && a.AllReduce;
}

// handles may be created from several threads: the table is filled exactly once
inline Api& api()
{
    static Api a;
    static std::once_flag once;
    std::call_once(once, load, std::ref(a));
    return a;
}

}  // namespace ncclx
