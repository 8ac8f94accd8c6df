// nccl_dl.h -- NCCL bound at run time with dlopen, so that libd3q19b200.so has no link-time
// NCCL dependency (single-GPU runs never touch it) and so that, inside a process that already
// loaded torch's bundled libnccl.so.2, the same library instance is reused.
// Only the handful of entry points the halo exchange and the scalar reductions need.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <mutex>

namespace d3q {

struct NcclUniqueId { char internal[128]; };            // nccl.h: NCCL_UNIQUE_ID_BYTES = 128
typedef struct ncclComm *NcclComm;
enum { NCCL_INT64 = 4, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8 };   // ncclDataType_t
enum { NCCL_SUM = 0, NCCL_MAX = 2 };                          // ncclRedOp_t

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

    bool ready = false;

    // returns nullptr on success, else a message.  Several host threads may create handles at the same time (one rank
    // per thread, host/channel_driver.cpp --ranks): the table is filled once, under a lock, and published last.
    const char *load() {
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        if (ready) return nullptr;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (lib) break;
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!lib) return "cannot dlopen libnccl.so.2";
#define D3Q_NCCL_SYM(field, name)                                     \
    *(void **)(&field) = dlsym(lib, name);                            \
    if (!field) return "libnccl is missing " name;
        D3Q_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        D3Q_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        D3Q_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        D3Q_NCCL_SYM(Send, "ncclSend")
        D3Q_NCCL_SYM(Recv, "ncclRecv")
        D3Q_NCCL_SYM(AllReduce, "ncclAllReduce")
        D3Q_NCCL_SYM(AllGather, "ncclAllGather")
        D3Q_NCCL_SYM(GroupStart, "ncclGroupStart")
        D3Q_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        D3Q_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef D3Q_NCCL_SYM
        ready = true;
        return nullptr;
    }
};

inline NcclApi &nccl_api() {
    static NcclApi api;
    return api;
}

}  // namespace d3q
