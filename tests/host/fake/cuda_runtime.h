// TEST INFRASTRUCTURE ONLY -- not CUDA, not a CPU fallback of the product.
//
// The smallest CUDA vocabulary that lets g++ compile d3q19-single-phase_b200/csrc/kernels.cuh as
// ordinary C++ (tests/host/kernels_host.cpp puts this directory first on the include path, so the
// `#include <cuda_runtime.h>` of kernels.cuh finds this file).  A "launch" is a loop nest over blocks
// and threads in ONE host thread, so only kernels whose threads do not cooperate can be run this way
// (the step, macrovar, AoS gather/scatter, face pack/unpack/put, vortcalc, FORCINGP, init, profile
// kernels).  __syncthreads is a no-op -- enough for the halo kernels, whose barriers only order a block's
// remote stores before its flag update -- and shuffles trap: kernels that reduce across a block (k_diag,
// k_rho_partial, block_max_to) are NOT run here.  Compile with -ffp-contract=off: R's + - * are then the same IEEE
// operations as __dadd_rn/__dsub_rn/__dmul_rn on the device.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__
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

static inline void hs_trap(const char *what) {
    std::fprintf(stderr, "host kernel harness: %s is not emulated\n", what);
    std::abort();
}
static inline void __syncthreads() {}      // threads of a block run one after the other: see the header comment
static inline void __threadfence_system() {}
static inline void __threadfence() {}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { hs_trap("__shfl_xor_sync"); return v; }
static inline long long __double_as_longlong(double v) { long long b; std::memcpy(&b, &v, 8); return b; }
static inline double __longlong_as_double(long long b) { double v; std::memcpy(&v, &b, 8); return v; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }

// launch<<<grid, block>>>: blocks in x-fastest order like the hardware's usual dispatch, threads in order
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
