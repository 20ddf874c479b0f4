// nccl_dyn.h -- NCCL bound at run time (dlopen), only when a handle is created with n_ranks > 1.
//
// libmgb200.so therefore has no link-time dependency on NCCL: single-GPU users never load it, and
// inside a process that already loaded a libnccl.so.2 (e.g. the one bundled with PyTorch) the same
// copy is reused instead of a second one with clashing symbols.  Only the stable core of the NCCL
// ABI is used (unique id, comm init/destroy, send/recv, all-reduce, groups).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <string>

namespace mgb {

struct NcclUniqueId { char internal[128]; };      // NCCL_UNIQUE_ID_BYTES
typedef struct ncclComm *NcclComm;
enum { kNcclSuccess = 0, kNcclSum = 0, kNcclMax = 2, kNcclUint8 = 1, kNcclUint64 = 5, kNcclFloat64 = 8 };

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;

    bool load()
    {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { error = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
#define MGB_NCCL_SYM(field, name)                                                   \
        field = reinterpret_cast<decltype(field)>(dlsym(lib, name));                \
        if (!field) { error = std::string("NCCL symbol missing: ") + name; return false; }
        MGB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        MGB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        MGB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        MGB_NCCL_SYM(Send, "ncclSend")
        MGB_NCCL_SYM(Recv, "ncclRecv")
        MGB_NCCL_SYM(AllReduce, "ncclAllReduce")
        MGB_NCCL_SYM(AllGather, "ncclAllGather")
        MGB_NCCL_SYM(GroupStart, "ncclGroupStart")
        MGB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        MGB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef MGB_NCCL_SYM
        return true;
    }
};

inline NcclApi &nccl()
{
    static NcclApi api;
    return api;
}

}  // namespace mgb
