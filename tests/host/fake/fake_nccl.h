// TEST INFRASTRUCTURE ONLY -- stands in for csrc/nccl_dl.h in the host-sim build (tests/host/make_hostsim.py): the same
// d3q::NcclApi interface, but every rank is a THREAD of this process and a message is a memcpy through a mailbox.
// Send is buffered and completes at once; Recv blocks until the matching message (same communicator, same source, FIFO)
// has been posted, and inside a group it is deferred to GroupEnd like NCCL does; AllReduce / AllGather meet at a
// barrier and reduce in rank order.  Stream arguments are ignored (the fake device is synchronous).
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <vector>

namespace d3q {

struct NcclUniqueId { char internal[128]; };
enum { NCCL_INT64 = 4, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_MAX = 2 };

struct FakeGroup {
    int nranks = 0;
    std::mutex mu;
    std::condition_variable cv;
    std::map<std::pair<int, int>, std::deque<std::vector<char>>> box;     // (src, dst) -> messages
    // collectives: everybody deposits, the last arrival opens the gate, everybody reads, the last reader resets
    std::vector<std::vector<char>> slot;
    int arrived = 0, readers = 0;
    unsigned gen = 0;
};
struct ncclComm { std::shared_ptr<FakeGroup> g; int rank; };
typedef ncclComm *NcclComm;

struct FakeRegistry {
    std::mutex mu;
    std::map<std::string, std::weak_ptr<FakeGroup>> groups;
};
inline FakeRegistry &fake_registry() { static FakeRegistry r; return r; }

struct PendingRecv { void *buf; size_t bytes; int peer; NcclComm comm; };
inline int &fake_group_depth() { static thread_local int d = 0; return d; }
inline std::vector<PendingRecv> &fake_pending() { static thread_local std::vector<PendingRecv> p; return p; }

inline size_t fake_dtype_size(int) { return 8; }      // only 64-bit types are used

inline int fake_do_recv(const PendingRecv &r) {
    FakeGroup &g = *r.comm->g;
    std::unique_lock<std::mutex> lk(g.mu);
    auto key = std::make_pair(r.peer, r.comm->rank);
    g.cv.wait(lk, [&] { return !g.box[key].empty(); });
    std::vector<char> m = std::move(g.box[key].front());
    g.box[key].pop_front();
    if (m.size() != r.bytes) return 5;                 // mismatched send/recv sizes
    std::memcpy(r.buf, m.data(), r.bytes);
    return 0;
}

template <class T>
inline void fake_reduce(T *out, const std::vector<std::vector<char>> &slot, size_t count, int op) {
    for (size_t i = 0; i < count; ++i) {
        T acc;
        std::memcpy(&acc, slot[0].data() + i * sizeof(T), sizeof(T));
        for (size_t r = 1; r < slot.size(); ++r) {
            T v;
            std::memcpy(&v, slot[r].data() + i * sizeof(T), sizeof(T));
            acc = (op == NCCL_SUM) ? (T)(acc + v) : (v > acc ? v : acc);
        }
        out[i] = acc;
    }
}

// mode 0: all-reduce, 1: all-gather
inline int fake_collective(int mode, const void *send, void *recv, size_t count, int dtype, int op, NcclComm c) {
    FakeGroup &g = *c->g;
    const size_t bytes = count * fake_dtype_size(dtype);
    std::unique_lock<std::mutex> lk(g.mu);
    g.cv.wait(lk, [&] { return g.readers == 0; });                 // the previous collective has been read by everybody
    if (g.slot.size() != (size_t)g.nranks) g.slot.assign(g.nranks, std::vector<char>());
    g.slot[c->rank].assign((const char *)send, (const char *)send + bytes);
    const unsigned my_gen = g.gen;
    if (++g.arrived == g.nranks) { g.arrived = 0; g.readers = g.nranks; ++g.gen; g.cv.notify_all(); }
    else g.cv.wait(lk, [&] { return g.gen != my_gen; });
    if (mode == 0) {
        if (dtype == NCCL_FLOAT64) fake_reduce((double *)recv, g.slot, count, op);
        else if (dtype == NCCL_INT64) fake_reduce((long long *)recv, g.slot, count, op);
        else fake_reduce((unsigned long long *)recv, g.slot, count, op);
    } else {
        for (int r = 0; r < g.nranks; ++r) std::memcpy((char *)recv + (size_t)r * bytes, g.slot[r].data(), bytes);
    }
    if (--g.readers == 0) g.cv.notify_all();
    return 0;
}

struct NcclConfig218 { int maxCTAs; };
inline NcclConfig218 nccl_config_max_ctas(int max_ctas) { return NcclConfig218{max_ctas}; }

struct NcclApi {
    const char *load() { return nullptr; }
    static int CommInitRankConfig(NcclComm *out, int nranks, NcclUniqueId id, int rank, NcclConfig218 *c) {
        if (!c || c->maxCTAs < 1) return 4;               // the library always asks for a positive CTA budget
        return CommInitRank(out, nranks, id, rank);
    }
    static int GetUniqueId(NcclUniqueId *id) {
        static std::mutex mu;
        static std::mt19937_64 rng(12345);
        std::lock_guard<std::mutex> lk(mu);
        for (int i = 0; i < 128; i += 8) { unsigned long long v = rng(); std::memcpy(id->internal + i, &v, 8); }
        return 0;
    }
    static int CommInitRank(NcclComm *out, int nranks, NcclUniqueId id, int rank) {
        FakeRegistry &reg = fake_registry();
        std::lock_guard<std::mutex> lk(reg.mu);
        const std::string key(id.internal, 128);
        std::shared_ptr<FakeGroup> g = reg.groups[key].lock();
        if (!g) { g = std::make_shared<FakeGroup>(); g->nranks = nranks; reg.groups[key] = g; }
        if (g->nranks != nranks) return 4;
        *out = new ncclComm{g, rank};
        return 0;
    }
    static int CommDestroy(NcclComm c) { delete c; return 0; }
    static int Send(const void *buf, size_t count, int dtype, int peer, NcclComm c, cudaStream_t) {
        FakeGroup &g = *c->g;
        const size_t bytes = count * fake_dtype_size(dtype);
        std::lock_guard<std::mutex> lk(g.mu);
        g.box[std::make_pair(c->rank, peer)].emplace_back((const char *)buf, (const char *)buf + bytes);
        g.cv.notify_all();
        return 0;
    }
    static int Recv(void *buf, size_t count, int dtype, int peer, NcclComm c, cudaStream_t) {
        PendingRecv r{buf, count * fake_dtype_size(dtype), peer, c};
        if (fake_group_depth() > 0) { fake_pending().push_back(r); return 0; }
        return fake_do_recv(r);
    }
    static int AllReduce(const void *s, void *r, size_t count, int dtype, int op, NcclComm c, cudaStream_t) {
        return fake_collective(0, s, r, count, dtype, op, c);
    }
    static int AllGather(const void *s, void *r, size_t count, int dtype, NcclComm c, cudaStream_t) {
        return fake_collective(1, s, r, count, dtype, 0, c);
    }
    static int GroupStart() { ++fake_group_depth(); return 0; }
    static int GroupEnd() {
        if (--fake_group_depth() > 0) return 0;
        std::vector<PendingRecv> todo;
        todo.swap(fake_pending());
        for (const PendingRecv &r : todo) { const int e = fake_do_recv(r); if (e) return e; }
        return 0;
    }
    static const char *GetErrorString(int e) { return e == 5 ? "fake NCCL: send/recv size mismatch" : "fake NCCL error"; }
};

inline NcclApi &nccl_api() {
    static NcclApi api;
    return api;
}

}  // namespace d3q
