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

// ncclConfig_t as of NCCL 2.18 (nccl.h: size, magic, version, then the user fields; later versions append fields and
// accept an older, shorter struct by its `size` and `version`).  Only maxCTAs is set here: the z faces are a few MB, and
// every CTA of NCCL's send/recv kernel occupies an SM next to the interior step kernel.
struct NcclConfig218 {
    size_t size;
    unsigned int magic;
    unsigned int version;
    int blocking, cgaClusterSize, minCTAs, maxCTAs;
    const char *netName;
    int splitShare;
};
inline NcclConfig218 nccl_config_max_ctas(int max_ctas) {
    const int undef = -2147483647 - 1;          // NCCL_CONFIG_UNDEF_INT
    NcclConfig218 c;
    c.size = sizeof(NcclConfig218); c.magic = 0xcafebeefu; c.version = 21800u;
    c.blocking = undef; c.cgaClusterSize = undef; c.minCTAs = undef; c.maxCTAs = max_ctas; c.netName = nullptr; c.splitShare = undef;
    return c;
}

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommInitRankConfig)(NcclComm *, int, NcclUniqueId, int, NcclConfig218 *) = nullptr;   // NCCL >= 2.14, optional
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
        *(void **)(&CommInitRankConfig) = dlsym(lib, "ncclCommInitRankConfig");
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
