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
#define __shared__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_ { unsigned x, y, z; };
static uint3_ threadIdx, blockIdx;
static dim3 blockDim, gridDim;

// ---- cooperative blocks: every thread of a block is a fiber (ucontext), barriers and shuffles yield ----------
namespace hs {
struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = true;
};
struct Barrier { int arrived = 0; unsigned gen = 0; };
static ucontext_t main_ctx;
static std::vector<Fiber> fibers;
static std::function<void()> body;
static int cur = -1;                      // running fiber, -1 = plain sequential launch
static Barrier block_bar, warp_bar[32];
static unsigned long long xch[1024];      // shuffle / vote exchange slots, one per thread

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
}  // namespace hs

static inline void __syncthreads() { hs::barrier(hs::block_bar, 0, (int)blockDim.x); }
static inline void __threadfence_system() {}
static inline void __threadfence() {}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int o) { return hs::warp_exchange(v, (hs::cur & 31) ^ o); }
template <class T> static inline T __shfl_sync(unsigned, T v, int lane) { return hs::warp_exchange(v, lane & 31); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int o) {
    const int lane = hs::cur & 31;
    const T r = hs::warp_exchange(v, lane >= o ? lane - o : lane);
    return lane >= o ? r : v;
}
static inline int __all_sync(unsigned, int pred) {
    int all = 1;
    for (int l = 0; l < 32; ++l) all &= hs::warp_exchange(pred ? 1 : 0, l);      // 32 rounds: simple, and only in tests
    return all;
}
static inline int __reduce_max_sync(unsigned, int v) {
    int m = v;
    for (int l = 0; l < 32; ++l) { const int o = hs::warp_exchange(v, l); m = o > m ? o : m; }
    return m;
}
static inline long long __double_as_longlong(double v) { long long b; std::memcpy(&b, &v, 8); return b; }
static inline double __longlong_as_double(long long b) { double v; std::memcpy(&v, &b, 8); return v; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
// one host thread runs every fiber: plain read-modify-write is atomic
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicCAS(T *p, T expect, T v) { T o = *p; if (o == expect) *p = v; return o; }

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
