// TEST INFRASTRUCTURE ONLY -- not CUDA, not a CPU fallback of the product.
//
// The smallest CUDA vocabulary that lets g++ compile d3q19-single-phase_b200/csrc/kernels.cuh as
// ordinary C++ (tests/host/kernels_host.cpp puts this directory first on the include path, so the
// `#include <cuda_runtime.h>` of kernels.cuh finds this file).  Everything runs in ONE host thread.
//   hs_launch       a loop nest over blocks and threads, for kernels whose threads do not cooperate (the step,
//                   macrovar, AoS gather/scatter, face pack/unpack/put, vortcalc, FORCINGP, init, profile kernels);
//                   __syncthreads is a no-op there -- enough for the halo kernels, whose barriers only order a
//                   block's remote stores before its flag update.
//   hs_launch_coop  the threads of a block are fibers (ucontext); __syncthreads, warp shuffles and votes are real
//                   rendezvous points, `__shared__` is a static, atomics are plain read-modify-writes.  For the
//                   reductions (k_diag, k_rho_partial, block_max_to) and the particle kernels (scan, warp-reduced
//                   force atomics).
// Compile with -ffp-contract=off: R's + - * are then the same IEEE operations as __dadd_rn/__dsub_rn/__dmul_rn
// on the device.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static thread_local      /* one block at a time per host thread */

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_ { unsigned x, y, z; };
static thread_local uint3_ threadIdx, blockIdx;       // thread_local: in the multi-rank host-sim every rank is a thread
static thread_local dim3 blockDim, gridDim;

// ---- cooperative blocks: every thread of a block is a fiber (ucontext), barriers and shuffles yield ----------
namespace hs {
struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = true;
};
struct Barrier { int arrived = 0; unsigned gen = 0; };
static thread_local ucontext_t main_ctx;
static thread_local std::vector<Fiber> fibers;
static thread_local std::function<void()> body;
static thread_local int cur = -1;         // running fiber, -1 = plain sequential launch
static thread_local Barrier block_bar, warp_bar[32];
static thread_local unsigned long long xch[1024];      // shuffle / vote exchange slots, one per thread

static inline void trap(const char *what) {
    std::fprintf(stderr, "host kernel harness: %s\n", what);
    std::abort();
}
static inline void yield() { swapcontext(&fibers[cur].ctx, &main_ctx); threadIdx.x = (unsigned)cur; }
static inline int alive(int first, int n) {
    int a = 0;
    for (int t = first; t < first + n && t < (int)blockDim.x; ++t) a += fibers[t].done ? 0 : 1;
    return a;
}
// all live threads of [first, first+n) meet here; threads that returned from the kernel do not count
static inline void barrier(Barrier &b, int first, int n) {
    if (cur < 0) return;                  // sequential launch: see the header comment
    const unsigned g = b.gen;
    ++b.arrived;
    while (b.gen == g) {
        if (b.arrived >= alive(first, n)) { b.arrived = 0; ++b.gen; break; }
        yield();
    }
}
static void trampoline() {
    body();
    fibers[cur].done = true;
    swapcontext(&fibers[cur].ctx, &main_ctx);
}
static inline void run_block(unsigned nthreads) {
    if (fibers.size() < nthreads) fibers.resize(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber &f = fibers[t];
        if (f.stack.empty()) f.stack.resize(512 * 1024);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.data();
        f.ctx.uc_stack.ss_size = f.stack.size();
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, trampoline, 0);
        f.done = false;
    }
    block_bar = Barrier();
    for (Barrier &w : warp_bar) w = Barrier();
    for (;;) {
        bool any = false;
        for (unsigned t = 0; t < nthreads; ++t) {
            if (fibers[t].done) continue;
            any = true;
            cur = (int)t;
            threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
            swapcontext(&main_ctx, &fibers[t].ctx);
        }
        if (!any) break;
    }
    cur = -1;
}
template <class T> static inline unsigned long long bits_of(T v) { unsigned long long b = 0; std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T from_bits(unsigned long long b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
// value of lane `src` of the calling thread's warp (all 32 lanes take part)
template <class T> static inline T warp_exchange(T v, int src_lane) {
    if (cur < 0) trap("warp shuffle in a sequential launch (use hs_launch_coop)");
    const int w = cur >> 5, first = cur & ~31;
    xch[cur] = bits_of(v);
    barrier(warp_bar[w], first, 32);
    const unsigned long long r = xch[first + src_lane];
    barrier(warp_bar[w], first, 32);
    return from_bits<T>(r);
}
// one rendezvous for a whole-warp vote: every lane deposits its predicate, then reads all 32
static inline unsigned warp_ballot(int pred) {
    if (cur < 0) trap("warp vote in a sequential launch (use hs_launch_coop)");
    const int w = cur >> 5, first = cur & ~31;
    xch[cur] = pred ? 1ull : 0ull;
    barrier(warp_bar[w], first, 32);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (unsigned)(xch[first + l] & 1ull) << l;
    barrier(warp_bar[w], first, 32);
    return m;
}
}  // namespace hs

static inline void __syncthreads() { hs::barrier(hs::block_bar, 0, (int)blockDim.x); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int o) { return hs::warp_exchange(v, (hs::cur & 31) ^ o); }
template <class T> static inline T __shfl_sync(unsigned, T v, int lane) { return hs::warp_exchange(v, lane & 31); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int o) {
    const int lane = hs::cur & 31;
    const T r = hs::warp_exchange(v, lane >= o ? lane - o : lane);
    return lane >= o ? r : v;
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, int o) {
    const int lane = hs::cur & 31;
    const T r = hs::warp_exchange(v, lane + o < 32 ? lane + o : lane);
    return lane + o < 32 ? r : v;
}
struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
// votes: one rendezvous (hs::warp_ballot); the maximum as an xor butterfly
static inline unsigned __ballot_sync(unsigned, int pred) { return hs::warp_ballot(pred); }
static inline int __all_sync(unsigned, int pred) { return __ballot_sync(0xffffffffu, pred) == 0xffffffffu; }
static inline int __any_sync(unsigned, int pred) { return __ballot_sync(0xffffffffu, pred) != 0u; }
static inline int __reduce_max_sync(unsigned, int v) {
    int m = v;
    for (int o = 16; o > 0; o >>= 1) { const int t = hs::warp_exchange(m, (hs::cur & 31) ^ o); m = t > m ? t : m; }
    return m;
}
static inline long long __double_as_longlong(double v) { long long b; std::memcpy(&b, &v, 8); return b; }
static inline double __longlong_as_double(long long b) { double v; std::memcpy(&v, &b, 8); return v; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
// the fibers of one rank share a host thread, so a plain read-modify-write would do there; the flag words of the
// peer-memory halo are touched by the threads of OTHER ranks in the multi-rank host-sim, hence real atomics
template <class T> static inline T atomicCAS(T *p, T expect, T v) { __atomic_compare_exchange_n(p, &expect, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return expect; }
template <class T> static inline T atomicAdd(T *p, T v) {
    T o = *p;
    for (;;) { const T seen = atomicCAS(p, o, (T)(o + v)); if (seen == o) return o; o = seen; }
}
static inline double atomicAdd(double *p, double v) {
    unsigned long long *q = (unsigned long long *)p, o = *q;
    for (;;) {
        double d; std::memcpy(&d, &o, 8); d += v;
        unsigned long long n; std::memcpy(&n, &d, 8);
        const unsigned long long seen = atomicCAS(q, o, n);
        if (seen == o) { double r; std::memcpy(&r, &o, 8); return r; }
        o = seen;
    }
}
template <class T> static inline T atomicMax(T *p, T v) {
    T o = *p;
    while (v > o) { const T seen = atomicCAS(p, o, v); if (seen == o) break; o = seen; }
    return o;
}

// launch<<<grid, block>>> for kernels whose threads do not cooperate: blocks in x-fastest order, threads in order
template <class K, class... Args>
static void hs_launch(dim3 grid, unsigned block, K kern, const Args &... args) {
    gridDim = grid;
    blockDim = dim3(block, 1, 1);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                for (unsigned t = 0; t < block; ++t) {
                    threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
                    kern(args...);
                }
            }
}
// the same for kernels that use __syncthreads, shared memory or warp shuffles: the threads of a block are fibers
template <class K, class... Args>
static void hs_launch_coop(dim3 grid, unsigned block, K kern, const Args &... args) {
    gridDim = grid;
    blockDim = dim3(block, 1, 1);
    hs::body = [&]() { kern(args...); };
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                hs::run_block(block);
            }
}

// ---- HS_FULL_RUNTIME: enough of the CUDA runtime API for csrc/d3q19_api.cu itself --------------------------------------
// tests/host/make_hostsim.py rewrites every `kernel<<<grid, block, smem, stream>>>(args)` of d3q19_api.cu into
// hs_dispatch("kernel", grid, block, kernel, args) and compiles the result with g++ into tests/host/_gen/ (git-ignored):
// the WHOLE C-ABI -- handle, transfers, step orchestration, shim state machine -- then runs on the build box, one
// synchronous "device" in host memory.  Streams and events are no-ops (everything completes at once), so this checks
// what the library computes and in which state it leaves its arrays, not the ordering between streams (that is
// tests/test_halo_schedule_model.py's job).  Still TEST INFRASTRUCTURE: nothing in the package builds or finds it.
#ifdef HS_FULL_RUNTIME
#include <chrono>
#include <string>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorFake = 1 };
typedef struct hs_stream_ *cudaStream_t;
struct hs_event_ { double t; };
typedef hs_event_ *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostRegisterDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "host-sim error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = (size_t)180 << 30; return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) {
    // poison: device memory is not zero-initialised either
    *p = (T *)std::malloc(n ? n : 1);
    if (!*p) return cudaErrorFake;
    std::memset((void *)*p, 0xA5, n);
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; ++r) std::memmove((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void *p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = (cudaStream_t)std::malloc(1); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline double hs_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new hs_event_{0.0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = hs_now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)((b->t - a->t) * 1e3); return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
// one process, one address space: a handle is the pointer
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { std::memset(h, 0, sizeof *h); std::memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }

// which launches need real block cooperation (everything else runs as a plain loop nest, ~10x faster)
template <class A> static inline bool hs_needs_coop(const A &) { return false; }
struct hs_none {};
template <class K, class A0, class... Rest>
static void hs_dispatch(const char *name, dim3 grid, unsigned block, K kern, const A0 &a0, const Rest &... rest) {
    static const char *coop[] = {"k_diag<", "k_rho_partial", "k_beads_links", "k_beads_lubmove", "k_beads_ibb"};
    bool c = hs_needs_coop(a0);              // found by ADL for d3q::StepParams (pre-relaxation: block maximum)
    for (const char *n : coop) c = c || !std::strncmp(name, n, std::strlen(n));
    if (c) hs_launch_coop(grid, block, kern, a0, rest...);
    else hs_launch(grid, block, kern, a0, rest...);
}
#endif   // HS_FULL_RUNTIME
