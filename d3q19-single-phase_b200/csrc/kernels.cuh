// kernels.cuh -- device kernels of the Channel-Flow time step for sm_100a.
//
// HBM layout: structure of arrays, x contiguous,
//     A[ip][zg][y][x],  x in [0,xp) (xp = lx rounded up to 16 doubles = 128 B),
//                       y in [0,ly), zg in [0,lz+2) with zg = 0 and lz+1 the ghost planes,
// so one population of one z plane ("face") is xp*ly contiguous doubles and a warp reads or
// writes 256 contiguous bytes per population.  rho/u/force arrays are [z][y][x] with the
// same pitch and no ghosts.
//
// A step is one pass: each thread owns one node, gathers its 19 populations according to
// the storage phase, collides in registers (collide.cuh) and scatters:
//     AB        read  g_i   = A[i][n - c_i]      (wall: A[opp i][n])   write B[i][n]
//     AA even   read  f_i   = A[i][n]                                  write A[opp i][n]
//     AA odd    read  f_i   = A[opp i][n - c_i]  (wall: A[i][n])       write f*_opp(i) to the
//                                                                      address f_i was read from
// 19 loads + 19 stores of 8 B per node = 304 B, nothing else in the main-loop instantiation.
// The reference gets the same post-streaming values with its sequential in-place "swap" sweep
// (collision.f90:224-264), which cannot be parallelised as written.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "collide.cuh"

namespace d3q {

enum ReadKind { READ_DIRECT = 0, READ_PULL_NAT = 1, READ_PULL_SWAP = 2 };
enum StepKind { STEP_AB = 0, STEP_AA_EVEN = 1, STEP_AA_ODD = 2 };

#ifndef D3Q_BLOCK_X
#define D3Q_BLOCK_X 128
#endif
// Resident CTAs per SM the step kernels are compiled for.  Measured on B200 (profiles/
// r01_variant_sweep.md, r01c_prefetch_sweep.md): AB and AA-even sit at the DRAM limit with 4 CTAs
// (<= 128 registers).  The AA odd step keeps 19 addresses live across the collision; without the L2
// prefetch it is faster with 3 CTAs and 155 registers (1.70 ms) than squeezed into 128 (1.78 ms,
// one load sinks below the collision), with the prefetch 4 CTAs win (1.55 ms vs 1.65 ms).
#ifndef D3Q_MIN_BLOCKS
#define D3Q_MIN_BLOCKS 4
#endif
#ifndef D3Q_MIN_BLOCKS_ODD
#define D3Q_MIN_BLOCKS_ODD 4
#endif
// the other AA odd instantiations (run-time macro mode / force field / solid mask, strict arithmetic,
// halo in peer memory) need more registers and would spill at 128
#ifndef D3Q_MIN_BLOCKS_ODD_WIDE
#define D3Q_MIN_BLOCKS_ODD_WIDE 3
#endif
constexpr int BLOCK_X = D3Q_BLOCK_X;

// cache policy of the population accesses (each address is read once and written once per step)
#ifndef D3Q_HINT
#define D3Q_HINT 0
#endif
// (Round 2 also tried accessors without the wall select for warps that hold no wall-adjacent node: 776 -> ~560 executed
//  instructions per node in the in-place odd step, +4 % while that step's prefetch was mis-aimed, -1 % once it was fixed
//  -- 128 registers with a spill against 124 without; removed.  profiles/r02c_variants.md, r02d_variants.md)
// 1: the in-place steps' L2 prefetch is issued by lane i for population i (2 instructions per warp), 0: by one lane per
// 128-byte line for all 19 populations
#ifndef D3Q_PF_LEAN
#define D3Q_PF_LEAN 1
#endif
#ifndef D3Q_PLAIN_WAIT           // 1: the plain step kernel can wait for the neighbours' flags itself (copy-engine transport)
#define D3Q_PLAIN_WAIT 1
#endif
#ifndef D3Q_ADDR                 // 0: wall handled by an offset select; 1: by a predicated second access
#define D3Q_ADDR 0
#endif
// TIMING EXPERIMENTS ONLY (tools/kernel_sweep.py; results are wrong): bit 0 = the AA odd step stores
// x-aligned, bit 1 = it also loads x-aligned.  Never set in the shipped library.
#ifndef D3Q_EXP
#define D3Q_EXP 0
#endif

// Software prefetch into L2: the step kernels are bound by memory latency at
// 12-16 resident warps per SM (ncu: long-scoreboard stalls, 0.39 eligible warps per cycle), so in the
// in-place (AA) steps one lane per 128-byte line asks L2 for the lines `pf_ahead` elements further
// down each population.  Every (population, node) element is read exactly once per step, so the
// lines of "my own node shifted ahead" are exactly what the blocks launched a little later will
// read; it holds no registers across the collision.  Measured on B200, 512x256x256
// (profiles/r01c_prefetch_sweep.md): ~128 blocks ahead is best (AA odd 1.706 -> 1.550 ms together with
// 4 CTAs/SM, AA even 1.519 -> 1.503 ms); >= 1 plane ahead thrashes L2; the two-array AB step gets
// SLOWER with any distance (1.512 -> 1.58 ms) and does not prefetch.
__device__ __forceinline__ void prefetch_l2(const void *p) {
#if defined(__CUDACC__)
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
#else
    (void)p;            // tests/host compiles this file with g++ (index logic checked on the GPU-less build box)
#endif
}

__device__ __forceinline__ double pop_load(const double *p) {
#if D3Q_HINT == 1
    return __ldcs(p);          // streaming: evict-first
#elif D3Q_HINT == 2
    return __ldcg(p);          // L2 only
#else
    return *p;
#endif
}
__device__ __forceinline__ void pop_store(double *p, double v) {
#if D3Q_HINT == 1
    __stcs(p, v);
#elif D3Q_HINT == 2
    __stcg(p, v);
#else
    *p = v;
#endif
}

struct Geom {
    int lx, ly, lz, xp;
    long long plane;      // xp*ly
    long long slab;       // plane*(lz+2): elements per population
    int zlo_src;          // ghosted plane index that acts as "z-1" of plane 1   (lz if 1 rank, else 0)
    int zhi_src;          // ghosted plane index that acts as "z+1" of plane lz  (1  if 1 rank, else lz+1)
};

// Halo over NVLink peer memory: the boundary planes of a step store their
// five outgoing populations straight into the neighbour GPU's array (cudaIpc-mapped) and raise a
// flag there; the neighbour's next step spins on that flag before it touches the planes.  No pack
// kernel, no NCCL call, no second stream: one launch per step.
struct Halo {
    double *peer_up, *peer_dn;          // the neighbours' array this step writes (population 0 base)
    long long slab_up, slab_dn;         // their elements per population
    int lz_dn;                          // thickness of the lower neighbour's slab
    unsigned int *wait_lo, *wait_hi;    // local flags raised by the lower / upper neighbour
    unsigned int *sig_up, *sig_dn;      // remote flags: the upper neighbour's wait_lo, the lower's wait_hi
    unsigned int *ctr;                  // two local block counters (plane 1, plane lz)
    unsigned int *err;                  // local watchdog word: the epoch whose flag never arrived, or 0
    unsigned long long timeout_ns;      // how long a boundary block spins before it gives up
    unsigned int epoch;                 // number of this halo step (1, 2, ...)
    unsigned int nblk_face;             // blocks per plane
};

__device__ __forceinline__ unsigned long long global_ns() {
#if defined(__CUDACC__)
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#else
    return 0ull;
#endif
}
// Spin until *flag >= want.  A neighbour that died (or a broken peer mapping) must not hang the GPU:
// after timeout_ns the waiter records `want` in *err and goes on -- the results are then void and
// d3q19_sync reports the failure.  Once *err is set every later waiter leaves at once.
__device__ __forceinline__ void halo_spin(const volatile unsigned int *flag, unsigned int want,
                                          unsigned int *err, unsigned long long timeout_ns) {
    if (*flag >= want) return;
    const unsigned long long t0 = global_ns();
    while (*flag < want) {
        if (*(volatile unsigned int *)err != 0u) return;
        if (global_ns() - t0 > timeout_ns) { atomicMax(err, want ? want : 1u); return; }
    }
}

struct StepParams {
    Geom g;
    double *A;            // source (and destination for AA)
    double *B;            // destination for AB
    int z0;               // first ghosted plane index handled by this launch
    int zstride;          // plane = z0 + blockIdx.z * zstride (boundary launch: planes 1 and lz in one grid)
    Mrt mrt;
    double Fx, Fy, Fz;    // uniform force (FORCING, collision.f90:522-524)
    double rho_shift;     // pending avedensity shift (collision.f90:505-511)
    long long pf_ahead;   // AA steps: L2 prefetch distance in elements (whole x-rows), 0 = off
    // generic instantiation only:
    int macro_mode;       // D3Q19_MACRO_*
    double *rho;          // device arrays [lz][ly][xp]
    const double *ux, *uy, *uz;
    const double *ffx, *ffy, *ffz;   // force field or nullptr
    const int32_t *solid;            // >0 solid, [lz][ly][xp], or nullptr
    unsigned long long *rhoerr_bits; // PRERELAX: max |rho_new - rho_old| as ordered bits
    Halo halo;                       // HALO instantiation only
};

// ---- neighbour addressing ------------------------------------------------------------------
// Offsets are element indices INSIDE one population (a "slab"); the population base
// A + slot*slab is warp-uniform.  IDX is uint32_t whenever a slab has < 2^32 elements (any
// realistic size: 34 GB per population), which halves the registers held across the collision
// in the AA odd step, where every address is used twice (load, then store).
template <class IDX>
struct NodeIdx {
    int x, y, zg;
    IDX n;                    // x + xp*(y + ly*zg)
    IDX row[3][3];            // x + xp*(y' + ly*z') for y' in (y-1,y,y+1), z' in (z-1,z,z+1), periodic
    bool wall_lo, wall_hi;    // x == 0 / x == lx-1
};

template <class IDX>
__device__ __forceinline__ NodeIdx<IDX> make_node(const Geom &g, int x, int y, int zg) {
    NodeIdx<IDX> k;
    k.x = x; k.y = y; k.zg = zg;
    const int ym = (y == 0) ? g.ly - 1 : y - 1;
    const int yp = (y == g.ly - 1) ? 0 : y + 1;
    const int zm = (zg == 1) ? g.zlo_src : zg - 1;
    const int zp = (zg == g.lz) ? g.zhi_src : zg + 1;
    const IDX oy[3] = {(IDX)ym * (IDX)g.xp, (IDX)y * (IDX)g.xp, (IDX)yp * (IDX)g.xp};
    const IDX oz[3] = {(IDX)zm * (IDX)g.plane, (IDX)zg * (IDX)g.plane, (IDX)zp * (IDX)g.plane};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) k.row[a][b] = (IDX)x + oy[a] + oz[b];
    k.n = k.row[1][1];
    k.wall_lo = (x == 0);
    k.wall_hi = (x == g.lx - 1);
    return k;
}

// Where direction I of node k is gathered from, for the three storage phases:
//   regular: population `slot_nb` at index n - c_i   (periodic y,z)
//   at a wall (x - c_ix outside): population `slot_wall` at the node itself (half-way
//   bounce-back, wall at rest -- the imove<1 .or. imove>lx branch of collision.f90:229-230)
template <int RK, int I>
struct Gather {
    static constexpr int cx = dir_cx(I), cy = dir_cy(I), cz = dir_cz(I), opp = dir_opp(I);
    static constexpr int slot_nb = (RK == READ_PULL_SWAP) ? opp : I;
    static constexpr int slot_wall = (RK == READ_PULL_SWAP) ? I : opp;
    static constexpr bool pulls = (RK != READ_DIRECT);
    static constexpr bool can_bounce = pulls && cx != 0;

    template <class IDX>
    static __device__ __forceinline__ bool at_wall(const NodeIdx<IDX> &k) {
        return can_bounce && ((cx > 0) ? k.wall_lo : k.wall_hi);
    }
    // index inside the population; the allocation is padded so that index -1 / +1 past the ends is mapped
    template <class IDX>
    static __device__ __forceinline__ IDX index_nb(const NodeIdx<IDX> &k) {
        return pulls ? (IDX)(k.row[1 - cy][1 - cz] - (IDX)cx) : k.n;
    }
    // one address per direction: the wall case is a select on the element offset, so every
    // direction costs exactly one LDG (and one STG in the AA odd step)
    template <class IDX>
    static __device__ __forceinline__ long long offset(const Geom &g, const NodeIdx<IDX> &k) {
        const long long reg = (long long)slot_nb * g.slab + (long long)index_nb(k);
        if (!can_bounce) return reg;
        return at_wall(k) ? (long long)slot_wall * g.slab + (long long)k.n : reg;
    }
    template <class IDX>
    static __device__ __forceinline__ double load(const double *A, const Geom &g, const NodeIdx<IDX> &k) {
#if D3Q_EXP & 2
        if (RK == READ_PULL_SWAP) return pop_load(A + offset(g, k) + cx);
#endif
#if D3Q_ADDR == 0
        return pop_load(A + offset(g, k));
#else
        double v = pop_load(A + (long long)slot_nb * g.slab + index_nb(k));
        if (can_bounce) {
            if (at_wall(k)) v = pop_load(A + (long long)slot_wall * g.slab + k.n);
        }
        return v;
#endif
    }
    // AA odd with the halo in peer memory: a value that would land in a z ghost plane is stored
    // into the neighbour's real plane instead (same slot, same row): ghost lz+1 -> the upper
    // neighbour's plane 1, ghost 0 -> the lower neighbour's plane lz_dn.
    template <class IDX>
    static __device__ __forceinline__ void store_back_halo(double *A, const Geom &g, const NodeIdx<IDX> &k, double v,
                                                           const Halo &h) {
        if (cz != 0 && !at_wall(k)) {
            if (cz < 0 && k.zg == g.lz) {          // pulled from z+1
                const long long o = (long long)index_nb(k) - (long long)(g.lz + 1) * g.plane + g.plane;
                pop_store(h.peer_up + (long long)slot_nb * h.slab_up + o, v);
                return;
            }
            if (cz > 0 && k.zg == 1) {             // pulled from z-1
                const long long o = (long long)index_nb(k) + (long long)h.lz_dn * g.plane;
                pop_store(h.peer_dn + (long long)slot_nb * h.slab_dn + o, v);
                return;
            }
        }
        store_back(A, g, k, v);
    }
    // AA odd: the post-collision value of direction opp(I) goes back to where f_I came from
    template <class IDX>
    static __device__ __forceinline__ void store_back(double *A, const Geom &g, const NodeIdx<IDX> &k, double v) {
#if D3Q_EXP & 1
        pop_store(A + offset(g, k) + cx, v); return;
#endif
#if D3Q_ADDR == 0
        pop_store(A + offset(g, k), v);
#else
        if (can_bounce) {
            if (at_wall(k)) { pop_store(A + (long long)slot_wall * g.slab + k.n, v); return; }
        }
        pop_store(A + (long long)slot_nb * g.slab + index_nb(k), v);
#endif
    }
};

template <int RK, class IDX>
__device__ __forceinline__ void gather19(const double *A, const Geom &g, const NodeIdx<IDX> &k, double (&f)[NPOP]) {
    static_for<NPOP>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        f[i] = Gather<RK, i>::load(A, g, k);
    });
}

// block maximum of a non-negative double, then one atomicMax per block: the bit pattern of a
// non-negative IEEE double is order-preserving as an unsigned integer
__device__ __forceinline__ void block_max_to(unsigned long long *dst, double v) {
    __shared__ unsigned long long smax[BLOCK_X / 32];
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
        b = t > b ? t : b;
    }
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int wi = 1; wi < BLOCK_X / 32; ++wi) b = smax[wi] > b ? smax[wi] : b;
        if (b) atomicMax(dst, b);
    }
}

// ---- the step kernel ---------------------------------------------------------------------------
// GENERIC = false: main loop, uniform force, no solids, moments in registers (304 B/node).
// GENERIC = true : run-time macro mode / force field / solid mask.
// HALO = true : z-slab run with the halo in peer memory; one launch covers the whole slab with the
//                two boundary planes first in block order (blockIdx.z 0 -> plane 1, 1 -> plane lz).
template <int SK, bool STRICT, bool GENERIC, class IDX, bool HALO = false>
__global__ void __launch_bounds__(BLOCK_X, SK != STEP_AA_ODD ? D3Q_MIN_BLOCKS
                                            : ((STRICT || GENERIC || HALO) ? D3Q_MIN_BLOCKS_ODD_WIDE : D3Q_MIN_BLOCKS_ODD))
k_step(const __grid_constant__ StepParams p) {
    const Geom &g = p.g;
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    double rhoerr = 0.0;
    const int zg_blk = HALO ? (blockIdx.z == 0 ? 1 : (blockIdx.z == 1 ? g.lz : (int)blockIdx.z))
                            : p.z0 + (int)blockIdx.z * p.zstride;
    // The planes next to a face read what the neighbour's previous step put here: their blocks wait for the neighbour's flag
    // (in lock step it arrived a whole interior sweep ago: one volatile load).  HALO: part of the fused transport.  Plain
    // instantiation: the boundary launch of the copy-engine transport passes the two flags (interior launches and single-GPU
    // runs pass nullptr) -- a one-thread wait kernel in front of every step would sit on the critical path instead.
    if (HALO ? (zg_blk == 1 || zg_blk == g.lz) : (D3Q_PLAIN_WAIT && p.halo.wait_lo != nullptr)) {
        if (threadIdx.x == 0) {
            const volatile unsigned int *fl = (zg_blk == 1) ? p.halo.wait_lo : p.halo.wait_hi;
            halo_spin(fl, p.halo.epoch - 1u, p.halo.err, p.halo.timeout_ns);
            __threadfence_system();
        }
        __syncthreads();
    }
    if (x < g.lx) {
        const NodeIdx<IDX> k = make_node<IDX>(g, x, blockIdx.y, zg_blk);
        constexpr int RK = (SK == STEP_AB) ? READ_PULL_NAT : (SK == STEP_AA_EVEN ? READ_DIRECT : READ_PULL_SWAP);
        double f[NPOP];
#if D3Q_PF_LEAN
        if (SK != STEP_AB && p.pf_ahead > 0) {
            // lane i < 19 asks for direction i: the two 128-byte lines that the warp pf_ahead elements further down will
            // READ for that direction (2 prefetch instructions per warp instead of 19 per line).  In the odd step that is
            // slot opp(i) at the node shifted by -c_i: with the shift left out, the ten c_z != 0 directions would ask for
            // lines that are read a whole z plane earlier or later -- harmless while a plane of all populations
            // (lx*ly*152 B) fits L2 next to everything else, 25 % extra DRAM reads on 1024x1024 planes (configs[3]),
            // where it does not (measured: odd step 1.93 ms vs 1.52 ms even, profiles/r02c_variants.jsonl).
            const int lane = threadIdx.x & 31;
            if (lane < NPOP) {
                const int slot = (RK == READ_PULL_SWAP) ? rt_opp(lane) : lane;
                const long long delta = (RK == READ_DIRECT) ? 0
                    : (long long)rt_cx(lane) + (long long)rt_cy(lane) * g.xp + (long long)rt_cz(lane) * g.plane;
                const long long ahead = (long long)k.n - lane + p.pf_ahead - delta;
                if (ahead >= 0 && ahead + 16 < g.slab) {      // stays inside the population
                    const double *q = p.A + (long long)slot * g.slab + ahead;
                    prefetch_l2(q);
                    prefetch_l2(q + 16);
                }
            }
        }
#else
        if (SK != STEP_AB && p.pf_ahead > 0 && (threadIdx.x & 15) == 0) {
            static_for<NPOP>([&](auto ic) {          // one lane per 128-byte line, all 19 directions (see the lean form)
                constexpr int i = decltype(ic)::value;
                constexpr int slot = (RK == READ_PULL_SWAP) ? dir_opp(i) : i;
                const long long delta = (RK == READ_DIRECT) ? 0
                    : (long long)dir_cx(i) + (long long)dir_cy(i) * g.xp + (long long)dir_cz(i) * g.plane;
                const long long ahead = (long long)k.n + p.pf_ahead - delta;
                if (ahead >= 0 && ahead < g.slab) prefetch_l2(p.A + (long long)slot * g.slab + ahead);
            });
        }
#endif
        gather19<RK>(p.A, g, k, f);

        double Fx = p.Fx, Fy = p.Fy, Fz = p.Fz;
        bool is_solid = false;
        if (GENERIC) {
            const long long m = x + (long long)g.xp * (k.y + (long long)g.ly * (k.zg - 1));
            if (p.ffx) { Fx = p.ffx[m]; Fy = p.ffy[m]; Fz = p.ffz[m]; }
            if (p.solid) is_solid = p.solid[m] > 0;
            if (!is_solid) {
                if (p.macro_mode == 0) {                      // D3Q19_MACRO_MAIN
                    if (STRICT) {
                        double r, a, b, c;
                        moments_strict(f, Fx, Fy, Fz, r, a, b, c);
                        collide_strict(f, (R(r) - R(p.rho_shift)).v, a, b, c, Fx, Fy, Fz, p.mrt);
                    } else {
                        collide_fast<true>(f, 0, 0, 0, 0, Fx, Fy, Fz, p.rho_shift, p.mrt);
                    }
                } else {
                    double r;
                    if (p.macro_mode == 1) {                  // D3Q19_MACRO_PRERELAX: fused rhoupdat
                        r = rho_index_order(f);
                        rhoerr = fabs(r - p.rho[m]);
                        p.rho[m] = r;
                    } else {
                        r = p.rho[m];
                    }
                    const double a = p.ux[m], b = p.uy[m], c = p.uz[m];
                    if (STRICT) collide_strict(f, r, a, b, c, Fx, Fy, Fz, p.mrt);
                    else collide_fast<false>(f, r, a, b, c, Fx, Fy, Fz, 0.0, p.mrt);
                }
            }
        } else {
            if (STRICT) {
                double r, a, b, c;
                moments_strict(f, Fx, Fy, Fz, r, a, b, c);
                collide_strict(f, r, a, b, c, Fx, Fy, Fz, p.mrt);
            } else {
                collide_fast<true>(f, 0, 0, 0, 0, Fx, Fy, Fz, 0.0, p.mrt);
            }
        }

        // in-plane offset of this node (y not decomposed: the same in every slab)
        const long long inplane = (long long)k.y * g.xp + x;
        if (SK == STEP_AB) {
            static_for<NPOP>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                pop_store(p.B + (long long)i * g.slab + k.n, f[i]);
                if (HALO && dir_cz(i) > 0 && zg_blk == g.lz)       // -> the upper neighbour's ghost plane 0
                    pop_store(p.halo.peer_up + (long long)i * p.halo.slab_up + inplane, f[i]);
                if (HALO && dir_cz(i) < 0 && zg_blk == 1)          // -> the lower neighbour's ghost plane lz_dn+1
                    pop_store(p.halo.peer_dn + (long long)i * p.halo.slab_dn + (long long)(p.halo.lz_dn + 1) * g.plane + inplane, f[i]);
            });
        } else if (SK == STEP_AA_EVEN) {
            static_for<NPOP>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                pop_store(p.A + (long long)dir_opp(i) * g.slab + k.n, f[i]);
                if (HALO && dir_cz(i) > 0 && zg_blk == g.lz)
                    pop_store(p.halo.peer_up + (long long)dir_opp(i) * p.halo.slab_up + inplane, f[i]);
                if (HALO && dir_cz(i) < 0 && zg_blk == 1)
                    pop_store(p.halo.peer_dn + (long long)dir_opp(i) * p.halo.slab_dn + (long long)(p.halo.lz_dn + 1) * g.plane + inplane, f[i]);
            });
        } else {
            static_for<NPOP>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                if (HALO) Gather<READ_PULL_SWAP, i>::store_back_halo(p.A, g, k, f[dir_opp(i)], p.halo);
                else Gather<READ_PULL_SWAP, i>::store_back(p.A, g, k, f[dir_opp(i)]);
            });
        }
    }
    if (HALO) {
        // last block of a boundary plane: everything it stored remotely is visible -> raise the neighbour's flag
        if (zg_blk == 1 || zg_blk == g.lz) {
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned int *c = p.halo.ctr + (zg_blk == g.lz ? 1 : 0);
                const unsigned int prev = atomicAdd(c, 1u);
                if (prev == p.halo.nblk_face - 1u) {
                    *c = 0u;
                    __threadfence_system();
                    *(volatile unsigned int *)(zg_blk == g.lz ? p.halo.sig_up : p.halo.sig_dn) = p.halo.epoch;
                }
            }
        }
    }
    if (GENERIC && p.macro_mode == 1 && p.rhoerr_bits) block_max_to(p.rhoerr_bits, rhoerr);
}

// 128-bit accesses (two x-adjacent nodes per thread, aligned double2 + lane shuffles for the c_x = +-1 populations)
// were built and measured in rounds 1-2 and REMOVED: 168 registers -> 3 CTAs/SM; on B200, 512x256x256, sustained:
// AB 1.823 ms vs 1.553 ms for the 64-bit kernel above, AA 1.692 vs 1.630 ms (profiles/r02_switch_decisions.md).  A warp
// of the 64-bit kernel already moves whole 128-byte lines and the DRAM bytes are the algorithmic ones; what the kernel
// needs is resident warps, not wider instructions.

// ---- initvel + initpop on the device (initial.f90:75-147, :19-46) -----------------------------------
// For fields too large to stage through the host (configs[3]: 150 GB of populations per GPU).  The
// velocity is a pure function of the GLOBAL node coordinates: log-law mean (initial.f90:104-115),
// the sinusoidal perturbation block (:119-144, amplitude A9; the reference ships A9 = 0) and a
// counter-based uniform noise (splitmix64 of (seed, component, global node); SURVEY.md 8(d)
// "synthetic inputs"), so every storage scheme can be filled without a halo exchange.
struct InitParams {
    Geom g;
    double *A;
    int nx, ny, nz, globalz;
    double ustar, ystar, A9, noise_amp, pi2;
    unsigned long long seed;
    int ivel;
    int unstream;         // 1: AB storage (post-collision, "un-streamed" canonical), 0: canonical
};

__host__ __device__ inline double splitmix_unit(unsigned long long seed, unsigned long long comp, unsigned long long node) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (3ull * node + comp + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;      // [-1, 1)
}

__device__ __forceinline__ void init_velocity(const InitParams &p, int ix, int iy, int izg, double &ux, double &uy, double &uz) {
    // ix in [1,nx], iy in [1,ny] (periodic), izg in [1,nz] global (periodic)
    ux = 0.0; uy = 0.0; uz = 0.0;
    if (p.ivel) {
        const int nxh = (p.nx + 1) / 2;
        const int i = ix <= nxh ? ix : p.nx + 1 - ix;          // mirrored about the centre plane (:107,:112)
        const double yplus = ((double)i - 0.5) / p.ystar;
        if (yplus < 10.8) uy = yplus * p.ustar;
        else uy = (log(yplus) / 0.41 + 5.0) * p.ustar;
        if (p.A9 != 0.0) {
            const double cc = 60.0;
            const double ccc1 = -(double)p.ny / p.pi2 / 1.0 / p.ystar * p.A9 * p.ustar / cc / cc;
            const double z9 = p.pi2 * ((double)izg - 0.5) / (double)p.nz;
            const double y9 = p.pi2 * ((double)iy - 0.5) / (double)p.ny;
            const double ccc9 = exp(-yplus / cc);
            uy += ccc1 * yplus * ccc9 * sin(y9 + z9);
            ux += p.A9 * p.ustar * (1. - ccc9 - yplus / cc * ccc9) * cos(y9 + z9);
        }
    }
    if (p.noise_amp != 0.0) {
        const unsigned long long node = (unsigned long long)(ix - 1) +
            (unsigned long long)p.nx * ((unsigned long long)(iy - 1) + (unsigned long long)p.ny * (unsigned long long)(izg - 1));
        ux += p.noise_amp * splitmix_unit(p.seed, 0, node);
        uy += p.noise_amp * splitmix_unit(p.seed, 1, node);
        uz += p.noise_amp * splitmix_unit(p.seed, 2, node);
    }
}

template <int I>
__device__ __forceinline__ double init_feq(double ux, double uy, double uz) {
    // initial.f90:26-44 with rho = 0
    const double usqr = 1.5 * (ux * ux + uy * uy + uz * uz);
    if (I == 0) return (1.0 / 3.0) * (0.0 - usqr);
    const double G = dir_cx(I) * ux + dir_cy(I) * uy + dir_cz(I) * uz;
    const double ww = I <= 6 ? 1.0 / 18.0 : 1.0 / 36.0;
    return ww * (0.0 + 3.0 * G + 4.5 * G * G - usqr);
}

__global__ void __launch_bounds__(BLOCK_X) k_init_channel(const __grid_constant__ InitParams p) {
    const Geom &g = p.g;
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= g.lx) return;
    const int y = blockIdx.y, zg = 1 + blockIdx.z;
    const long long n = x + (long long)g.xp * (y + (long long)g.ly * zg);
    const int ix = x + 1, iy = y + 1, iz = p.globalz + zg;          // global 1-based
    double ux, uy, uz;
    init_velocity(p, ix, iy, iz, ux, uy, uz);
    static_for<NPOP>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        double v;
        if (!p.unstream) {
            v = init_feq<i>(ux, uy, uz);
        } else {
            // AB storage holds g_i(x) = f_i(x + c_i); at a wall g_i(x) = f_opp(i)(x)
            constexpr int cx = dir_cx(i), cy = dir_cy(i), cz = dir_cz(i);
            const int jx = ix + cx;
            if (jx < 1 || jx > p.nx) {
                v = init_feq<dir_opp(i)>(ux, uy, uz);
            } else if (i == 0) {
                v = init_feq<0>(ux, uy, uz);
            } else {
                int jy = iy + cy, jz = iz + cz;
                jy = jy < 1 ? p.ny : (jy > p.ny ? 1 : jy);
                jz = jz < 1 ? p.nz : (jz > p.nz ? 1 : jz);
                double a, b, c;
                init_velocity(p, jx, jy, jz, a, b, c);
                v = init_feq<i>(a, b, c);
            }
        }
        p.A[(long long)i * g.slab + n] = v;
    });
}

// readers of the ghost planes (download, macrovar, probe, profiles) run after the neighbours' stores
__global__ void k_halo_wait(const volatile unsigned int *lo, const volatile unsigned int *hi, unsigned int epoch,
                            unsigned int *err, unsigned long long timeout_ns) {
    halo_spin(lo, epoch, err, timeout_ns);
    halo_spin(hi, epoch, err, timeout_ns);
    __threadfence_system();
}

// ---- macrovar (collision.f90:378-463) / rhoupdat (:469-480) -----------------------------------
struct MacroParams {
    Geom g;
    const double *A;
    double *rho, *ux, *uy, *uz;       // [lz][ly][xp]
    double Fx, Fy, Fz;
    const double *ffx, *ffy, *ffz;
    const int32_t *solid;             // ibnodes > 0
    const int32_t *isnodes;           // owning particle, 1-based
    const double *ypglb, *wp, *omgp;  // (3,npart)
    double rhopart;
    int ipart, ny, nz, globalz;
    int rho_only;                     // rhoupdat: all nodes, index order, rho only
    unsigned long long *rhoerr_bits;  // rhoupdat: max |rho_new - rho_old| as ordered bits (main.f90:79), or nullptr
};

template <int RK>
__global__ void __launch_bounds__(BLOCK_X) k_macro(const __grid_constant__ MacroParams p) {
    const Geom &g = p.g;
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (p.rho_only && p.rhoerr_bits) {
        // whole blocks stay alive for the block maximum
        double err = 0.0;
        if (x < g.lx) {
            const NodeIdx<unsigned long long> k = make_node<unsigned long long>(g, x, blockIdx.y, 1 + blockIdx.z);
            const long long m = x + (long long)g.xp * (k.y + (long long)g.ly * (k.zg - 1));
            double f[NPOP];
            gather19<RK>(p.A, g, k, f);
            const double r = rho_index_order(f);
            err = fabs(r - p.rho[m]);
            p.rho[m] = r;
        }
        block_max_to(p.rhoerr_bits, err);
        return;
    }
    if (x >= g.lx) return;
    const NodeIdx<unsigned long long> k = make_node<unsigned long long>(g, x, blockIdx.y, 1 + blockIdx.z);
    const long long m = x + (long long)g.xp * (k.y + (long long)g.ly * (k.zg - 1));
    double f[NPOP];
    gather19<RK>(p.A, g, k, f);
    if (p.rho_only) { p.rho[m] = rho_index_order(f); return; }
    const bool is_solid = p.solid && p.solid[m] > 0;
    if (!is_solid) {
        double Fx = p.Fx, Fy = p.Fy, Fz = p.Fz;
        if (p.ffx) { Fx = p.ffx[m]; Fy = p.ffy[m]; Fz = p.ffz[m]; }
        double r, a, b, c;
        moments_strict(f, Fx, Fy, Fz, r, a, b, c);
        p.rho[m] = r; p.ux[m] = a; p.uy[m] = b; p.uz[m] = c;
    } else if (p.ipart) {
        // rigid-body velocity of the owning particle, nearest periodic image in y,z (collision.f90:420-459)
        const int id = p.isnodes[m] - 1;
        const double xpnt = (double)(x + 1) - 0.5;
        const double ypnt = (double)(k.y + 1) - 0.5;                 // y is not decomposed: globaly = 0
        const double zpnt = (double)(k.zg) - 0.5 + (double)p.globalz;
        double xc = p.ypglb[3 * id], yc = p.ypglb[3 * id + 1], zc = p.ypglb[3 * id + 2];
        const double nyh = (double)(p.ny / 2), nzh = (double)(p.nz / 2);
        if ((yc - ypnt) > nyh) yc = yc - (double)p.ny;
        if ((yc - ypnt) < -nyh) yc = yc + (double)p.ny;
        if ((zc - zpnt) > nzh) zc = zc - (double)p.nz;
        if ((zc - zpnt) < -nzh) zc = zc + (double)p.nz;
        const double xx0 = xpnt - xc, yy0 = ypnt - yc, zz0 = zpnt - zc;
        const double w1 = p.wp[3 * id], w2 = p.wp[3 * id + 1], w3 = p.wp[3 * id + 2];
        const double o1 = p.omgp[3 * id], o2 = p.omgp[3 * id + 1], o3 = p.omgp[3 * id + 2];
        p.ux[m] = (R(w1) + (R(o2) * R(zz0) - R(o3) * R(yy0))).v;
        p.uy[m] = (R(w2) + (R(o3) * R(xx0) - R(o1) * R(zz0))).v;
        p.uz[m] = (R(w3) + (R(o1) * R(yy0) - R(o2) * R(xx0))).v;
        p.rho[m] = p.rhopart;
    }
}

// one node -> out[4] (probe, saveload.f90:4059-4100)
// (ffx,ffy,ffz: the force field when one is set -- the velocity is momentum + F/2 with the node's own force,
// collision.f90:415-417 -- else nullptr and the uniform force applies; same in k_profiles and k_diag)
template <int RK>
__global__ void k_probe(Geom g, const double *A, int x, int y, int zg, double Fx, double Fy, double Fz, const double *ffx,
                        const double *ffy, const double *ffz, double *out) {
    const NodeIdx<unsigned long long> k = make_node<unsigned long long>(g, x, y, zg);
    double f[NPOP];
    gather19<RK>(A, g, k, f);
    if (ffx) {
        const long long m = x + (long long)g.xp * (y + (long long)g.ly * (zg - 1));
        Fx = ffx[m]; Fy = ffy[m]; Fz = ffz[m];
    }
    moments_strict(f, Fx, Fy, Fz, out[0], out[1], out[2], out[3]);
}

// ---- vortcalc (saveload.f90:3929-4025) ---------------------------------------------------------------
// Vorticity of the device-resident ux,uy,uz ([lz][ly][xp], made by k_macro): central differences,
// one-sided at the channel walls (:3971-3980), periodic in y by index wrap (the reference's tmpu?L/R
// rows of exchng8 collapse for one rank in y), z neighbours across a slab face from the planes the
// caller exchanged (exchng8's z phase, :4039-4045): zlo = the lower neighbour's plane lz, zhi = the
// upper neighbour's plane 1, three fields of `plane` elements each; nullptr = one rank, periodic wrap.
// A solid node gets twice its particle's angular velocity (:4008-4019).  Same expression order as the
// reference without contraction: bit-identical to the oracle.
struct VortParams {
    Geom g;
    const double *ux, *uy, *uz;
    double *ox, *oy, *oz;
    const double *zlo, *zhi;
    const int32_t *solid, *isnodes;
    const double *omgp;
};
__global__ void __launch_bounds__(BLOCK_X) k_vortcalc(const __grid_constant__ VortParams p) {
    const Geom &g = p.g;
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= g.lx) return;
    const int y = blockIdx.y, z = blockIdx.z;                  // 0-based local
    const long long row = (long long)g.xp * (y + (long long)g.ly * z);
    const long long m = row + x;
    if (p.solid && !(p.solid[m] < 0)) {
        const int id = p.isnodes[m] - 1;
        p.ox[m] = (R(2.0) * R(p.omgp[3 * id])).v;
        p.oy[m] = (R(2.0) * R(p.omgp[3 * id + 1])).v;
        p.oz[m] = (R(2.0) * R(p.omgp[3 * id + 2])).v;
        return;
    }
    const int ym = (y == 0) ? g.ly - 1 : y - 1, yp = (y == g.ly - 1) ? 0 : y + 1;
    const long long rym = (long long)g.xp * (ym + (long long)g.ly * z) + x;
    const long long ryp = (long long)g.xp * (yp + (long long)g.ly * z) + x;
    const long long inpl = (long long)g.xp * y + x;
    // z neighbours: inside the slab, or the exchanged planes, or the periodic wrap of a single rank
    const double *uxm, *uym, *uxp, *uyp;
    if (z > 0) { uxm = p.ux + m - g.plane; uym = p.uy + m - g.plane; }
    else if (p.zlo) { uxm = p.zlo + inpl; uym = p.zlo + g.plane + inpl; }
    else { uxm = p.ux + (long long)(g.lz - 1) * g.plane + inpl; uym = p.uy + (long long)(g.lz - 1) * g.plane + inpl; }
    if (z < g.lz - 1) { uxp = p.ux + m + g.plane; uyp = p.uy + m + g.plane; }
    else if (p.zhi) { uxp = p.zhi + inpl; uyp = p.zhi + g.plane + inpl; }
    else { uxp = p.ux + inpl; uyp = p.uy + inpl; }
    R pwx, pvx;
    if (x == 0) {
        pwx = R(__ddiv_rn((R(3.0) * R(p.uz[m]) + R(p.uz[m + 1])).v, 3.0));
        pvx = R(__ddiv_rn((R(3.0) * R(p.uy[m]) + R(p.uy[m + 1])).v, 3.0));
    } else if (x == g.lx - 1) {
        pwx = R(__ddiv_rn(-(R(3.0) * R(p.uz[m]) + R(p.uz[m - 1])).v, 3.0));
        pvx = R(__ddiv_rn(-(R(3.0) * R(p.uy[m]) + R(p.uy[m - 1])).v, 3.0));
    } else {
        pwx = R(__ddiv_rn((R(p.uz[m + 1]) - R(p.uz[m - 1])).v, 2.0));
        pvx = R(__ddiv_rn((R(p.uy[m + 1]) - R(p.uy[m - 1])).v, 2.0));
    }
    const R pwy = R(__ddiv_rn((R(p.uz[ryp]) - R(p.uz[rym])).v, 2.0));
    const R puy = R(__ddiv_rn((R(p.ux[ryp]) - R(p.ux[rym])).v, 2.0));
    const R puz = R(__ddiv_rn((R(*uxp) - R(*uxm)).v, 2.0));
    const R pvz = R(__ddiv_rn((R(*uyp) - R(*uym)).v, 2.0));
    p.ox[m] = (pwy - pvz).v;
    p.oy[m] = (puz - pwx).v;
    p.oz[m] = (pvx - puy).v;
}

// ---- local strain rate from the non-equilibrium moments (sijstat00, saveload.f90:2031-2091) ----------------
// Sij*Sij of every fluid node from its 19 populations (canonical = post-streaming, whatever the storage phase) and
// the rho,u arrays macrovar left on the device: the second-order moments minus their equilibria, times their
// relaxation rates, ARE the strain-rate tensor up to constants (Yu et al. 2006) -- node-local, no differences, no
// halo.  The sums are collision_MRT's; the reference's expression order without contraction: bit-identical to the
// translated reference.  Solid nodes get 0 (the reference leaves its automatic array undefined there).
struct SijParams {
    Geom g;
    const double *A;
    const double *rho, *ux, *uy, *uz;   // [lz][ly][xp]
    const int32_t *solid;               // > 0 solid, or nullptr
    double s1, s9;
    double *sij2;                       // [lz][ly][xp]
};
template <int RK>
__global__ void __launch_bounds__(BLOCK_X) k_sijstat(const __grid_constant__ SijParams p) {
    const Geom &g = p.g;
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= g.lx) return;
    const int y = blockIdx.y, zg = 1 + blockIdx.z;
    const long long m = x + (long long)g.xp * (y + (long long)g.ly * (zg - 1));
    if (p.solid && p.solid[m] > 0) { p.sij2[m] = 0.0; return; }
    const NodeIdx<unsigned long long> k = make_node<unsigned long long>(g, x, y, zg);
    double f[NPOP];
    gather19<RK>(p.A, g, k, f);
    const R coef2(-11.0), coef3(8.0), coef5(2.0);                      // para.f90:144-147
    const R rho9(p.rho[m]), ux9(p.ux[m]), uy9(p.uy[m]), uz9(p.uz[m]);
    const R ux9s = ux9 * ux9, uy9s = uy9 * uy9, uz9s = uz9 * uz9;
    const R eqm1 = -(R(11.0) * rho9) + R(19.0) * ((ux9s + uy9s) + uz9s);
    const R eqm6 = (R(2.0) * ux9s - uy9s) - uz9s;
    const R eqm8 = uy9s - uz9s;
    const R eqm10 = ux9 * uy9, eqm11 = uy9 * uz9, eqm12 = ux9 * uz9;
    auto F = [&](int i) { return R(f[i]); };
    const R sum1 = ((((F(1) + F(2)) + F(3)) + F(4)) + F(5)) + F(6);
    const R sum2 = ((((((((((F(7) + F(8)) + F(9)) + F(10)) + F(11)) + F(12)) + F(13)) + F(14)) + F(15)) + F(16)) + F(17)) + F(18);
    const R sum6 = F(1) + F(2);
    const R sum7 = ((F(3) + F(4)) + F(5)) + F(6);
    const R sum8 = ((((((F(7) + F(8)) + F(9)) + F(10)) + F(11)) + F(12)) + F(13)) + F(14);
    const R sum9 = ((F(15) + F(16)) + F(17)) + F(18);
    const R sum10 = ((F(3) + F(4)) - F(5)) - F(6);
    const R sum11 = ((((((F(7) + F(8)) + F(9)) + F(10)) - F(11)) - F(12)) - F(13)) - F(14);
    const R evlm1 = (-(R(30.0) * F(0)) + coef2 * sum1) + coef3 * sum2;
    const R evlm6 = ((coef5 * sum6 - sum7) + sum8) - coef5 * sum9;
    const R evlm8 = sum10 + sum11;
    const R evlm10 = ((F(7) - F(8)) - F(9)) + F(10);
    const R evlm11 = ((F(15) - F(16)) - F(17)) + F(18);
    const R evlm12 = ((F(11) - F(12)) - F(13)) + F(14);
    const R s1(p.s1), s9(p.s9);
    const R neqm1 = s1 * (evlm1 - eqm1), neqm9 = s9 * (evlm6 - eqm6), neqm11 = s9 * (evlm8 - eqm8);
    const R neqm13 = s9 * (evlm10 - eqm10), neqm14 = s9 * (evlm11 - eqm11), neqm15 = s9 * (evlm12 - eqm12);
    const R Sxx = -R(__ddiv_rn((neqm1 + R(19.0) * neqm9).v, 38.0));
    const R Syy = -R(__ddiv_rn((R(2.0) * neqm1 - R(19.0) * (neqm9 - R(3.0) * neqm11)).v, 76.0));
    const R Szz = -R(__ddiv_rn((R(2.0) * neqm1 - R(19.0) * (neqm9 + R(3.0) * neqm11)).v, 76.0));
    const R Sxy = -(R(1.5) * neqm13), Syz = -(R(1.5) * neqm14), Szx = -(R(1.5) * neqm15);
    p.sij2[m] = (((Sxx * Sxx + Syy * Syy) + Szz * Szz) + R(2.0) * ((Sxy * Sxy + Syz * Syz) + Szx * Szx)).v;
}

// ---- FORCINGP (collision.f90:529-602) ------------------------------------------------------------------
// Time-dependent perturbation force in two near-wall x-bands on top of the uniform (0, force_in_y, 0):
// band 1 = nodes ixs0+1 .. ixs0+ihh, band 2 = nodes nx-ixs0-ihh+1 .. nx-ixs0 (1-based), ihh = lxh/2.
// The reference fills the three host arrays every call; here the device force field is written directly
// (no PCIe).  Amp0 (one sin of the step number) comes from the host; the per-node expressions keep the
// reference's order without contraction, so the only difference to the reference is the last-bit
// behaviour of the device sin/cos (tests: 1e-14 of the field's maximum).
struct ForcingpParams {
    int lx, ly, lz, xp, nx, ny, nz, globalz;
    int ihh, ixs0;
    double force_in_y, Amp0, beta9, gamma9, phase9, pi2;
    double *fx, *fy, *fz;
};
__global__ void __launch_bounds__(BLOCK_X) k_forcingp(const __grid_constant__ ForcingpParams p) {
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= p.lx) return;
    const int y = blockIdx.y, z = blockIdx.z;
    const long long m = x + (long long)p.xp * (y + (long long)p.ly * z);
    const int ig = x + 1, jj = y + 1, kk = z + 1 + p.globalz;          // global 1-based (y not decomposed)
    double fx = 0.0, fy = p.force_in_y, fz = 0.0;                      // :546-548
    // the second band is written after the first (:582-598), so it wins where the two overlap
    const int ixs2 = p.nx - p.ixs0 - p.ihh;
    int band = 0, i = 0;
    if (ig > ixs2 && ig <= ixs2 + p.ihh) { band = 2; i = ig - ixs2; }
    else if (ig > p.ixs0 && ig <= p.ixs0 + p.ihh) { band = 1; i = ig - p.ixs0; }
    if (band) {
        const R fiy(p.force_in_y), amp(p.Amp0), half(0.5), one(1.0);
        const double z9 = __ddiv_rn((R(p.pi2) * (R((double)kk) - half)).v, (double)p.nz);
        const double yfrac = __ddiv_rn((R((double)jj) - half).v, (double)p.ny);
        const double y9 = band == 1 ? __ddiv_rn((R(p.pi2) * (R((double)jj) - half)).v, (double)p.ny)
                                    : (R(p.pi2) * (R(yfrac) + R(p.phase9))).v;
        const double x9 = __ddiv_rn((R(p.pi2) * (R((double)i) - half)).v, (double)p.ihh);
        const double by = (R(p.beta9) * R(y9)).v, gz = (R(p.gamma9) * R(z9)).v;
        const R sx(sin(x9)), cxm(__dsub_rn(1.0, cos(x9))), sby(sin(by)), cby(cos(by)), sgz(sin(gz)), cgz(cos(gz));
        const R lead = band == 1 ? fiy : R(-p.force_in_y);
        // force_in_y*0.5*Amp0*real(ihh)*(1.-cos(x9))*cos(beta9*y9)*cos(gamma9*z9)
        fx = (((((lead * half) * amp) * R((double)p.ihh)) * cxm) * cby * cgz).v;
        // force_in_y*(1.0 -+ Amp0*real(ny)/beta9*sin(x9)*sin(beta9*y9)*cos(gamma9*z9))
        const R t = ((R(__ddiv_rn((amp * R((double)p.ny)).v, p.beta9)) * sx) * sby) * cgz;
        fy = (fiy * (band == 1 ? one - t : one + t)).v;
        // force_in_y*0.5*Amp0*real(nz)/gamma9*sin(x9)*cos(beta9*y9)*sin(gamma9*z9)
        fz = (((R(__ddiv_rn((((lead * half) * amp) * R((double)p.nz)).v, p.gamma9)) * sx) * cby) * sgz).v;
    }
    p.fx[m] = fx; p.fy[m] = fy; p.fz[m] = fz;
}

// ---- canonical AoS <-> device SoA (upload_f / download_f) ----------------------------------------
// aos holds planes [zg0, zg0+nz) of the host layout f(0:18,lx,ly,:) without pitch.
template <int RK>
__global__ void __launch_bounds__(BLOCK_X) k_gather_aos(Geom g, const double *A, double *aos, int zg0) {
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= g.lx) return;
    const NodeIdx<unsigned long long> k = make_node<unsigned long long>(g, x, blockIdx.y, zg0 + blockIdx.z);
    double f[NPOP];
    gather19<RK>(A, g, k, f);
    double *o = aos + (long long)NPOP * (x + (long long)g.lx * (k.y + (long long)g.ly * blockIdx.z));
#pragma unroll
    for (int i = 0; i < NPOP; ++i) o[i] = f[i];
}

__global__ void __launch_bounds__(BLOCK_X) k_scatter_aos(Geom g, double *A, const double *aos, int zg0) {
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= g.lx) return;
    const int y = blockIdx.y, zg = zg0 + blockIdx.z;
    const long long n = x + (long long)g.xp * (y + (long long)g.ly * zg);
    const double *s = aos + (long long)NPOP * (x + (long long)g.lx * (y + (long long)g.ly * blockIdx.z));
#pragma unroll
    for (int i = 0; i < NPOP; ++i) A[(long long)i * g.slab + n] = s[i];
}

// AB upload: canonical C (post-streaming) -> post-collision storage, the inverse permutation
// of streaming:  g_i(x) = f_i(x + c_i), at a wall g_i(x) = f_opp(i)(x).
__global__ void __launch_bounds__(BLOCK_X) k_unstream(Geom g, const double *Cn, double *A) {
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= g.lx) return;
    const NodeIdx<unsigned long long> k = make_node<unsigned long long>(g, x, blockIdx.y, 1 + blockIdx.z);
    static_for<NPOP>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int cx = dir_cx(i), cy = dir_cy(i), cz = dir_cz(i);
        const bool wall = (cx > 0) ? k.wall_hi : (cx < 0 ? k.wall_lo : false);
        const unsigned long long nb = k.row[1 + cy][1 + cz] + (unsigned long long)(long long)cx;
        A[(long long)i * g.slab + k.n] =
            wall ? Cn[(long long)dir_opp(i) * g.slab + k.n] : Cn[(long long)i * g.slab + nb];
    });
}

// pitched scalar field <-> host layout (lx,ly,lz)
__global__ void __launch_bounds__(BLOCK_X) k_field_pack(int lx, int xp, const double *dev, double *host_layout, int to_host) {
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= lx) return;
    const long long row = blockIdx.y + (long long)gridDim.y * blockIdx.z;
    if (to_host) host_layout[row * lx + x] = dev[row * xp + x];
}
__global__ void __launch_bounds__(BLOCK_X) k_field_unpack(int lx, int xp, double *dev, const double *host_layout) {
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= lx) return;
    const long long row = blockIdx.y + (long long)gridDim.y * blockIdx.z;
    dev[row * xp + x] = host_layout[row * lx + x];
}
__global__ void __launch_bounds__(BLOCK_X) k_field_unpack_i32(int lx, int xp, int32_t *dev, const int32_t *host_layout,
                                                              int src_ghosted, int ly) {
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= lx) return;
    const int y = blockIdx.y, z = blockIdx.z;
    const long long row = y + (long long)gridDim.y * z;
    long long s;
    if (src_ghosted)   // ibnodes(0:lx+1,0:ly+1,0:lz+1)
        s = (x + 1) + (long long)(lx + 2) * ((y + 1) + (long long)(ly + 2) * (z + 1));
    else
        s = row * lx + x;
    dev[row * xp + x] = host_layout[s];
}

// ---- z-face pack / unpack (collisionExchnge, collision.f90:337-370) ------------------------------
// buf[s][y][x] (pitch xp) <- A[slots[s]][plane][y][x]
struct FaceSlots { int s[5]; };
// both faces in one launch: blockIdx.z in [0,10), face = z / 5 (0: "up" data, 1: "dn" data)
struct FacePair {
    double *buf[2];
    int zg[2];
    FaceSlots slots[2];
};
__global__ void __launch_bounds__(BLOCK_X) k_face_pack(Geom g, const double *A, FacePair fp) {
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= g.xp) return;
    const int y = blockIdx.y, face = blockIdx.z / 5, s = blockIdx.z % 5;
    fp.buf[face][(long long)s * g.plane + (long long)y * g.xp + x] =
        A[(long long)fp.slots[face].s[s] * g.slab + (long long)fp.zg[face] * g.plane + (long long)y * g.xp + x];
}
// exclude_walls: populations with c_x = +1 skip x = 0, c_x = -1 skip x = lx-1 -- the slices
// 2:lx / 1:lx-1 of collision.f90:361-362,367-368 that keep the locally bounced value.
__global__ void __launch_bounds__(BLOCK_X) k_face_unpack(Geom g, double *A, FacePair fp, int exclude_walls) {
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= g.lx) return;
    const int y = blockIdx.y, face = blockIdx.z / 5, s = blockIdx.z % 5;
    const int slot = fp.slots[face].s[s];
    if (exclude_walls) {
        const int cx = (slot == 11 || slot == 13 || slot == 7 || slot == 9 || slot == 1) ? 1
                     : ((slot == 12 || slot == 14 || slot == 8 || slot == 10 || slot == 2) ? -1 : 0);
        if ((cx > 0 && x == 0) || (cx < 0 && x == g.lx - 1)) return;
    }
    A[(long long)slot * g.slab + (long long)fp.zg[face] * g.plane + (long long)y * g.xp + x] =
        fp.buf[face][(long long)s * g.plane + (long long)y * g.xp + x];
}

// ---- z-face "put" (halo transport by the copy engines) --------------------------------
// The faces are device-to-device copies over NVLink enqueued by the host (d3q19_api.cu launch_step_put); what is
// left for a kernel is to tell the neighbours that their planes are complete: one thread, after the copies in stream
// order, raises the upper neighbour's wait_lo and the lower neighbour's wait_hi to the step number (halo_spin protocol).
__host__ __device__ inline int dir_cx_rt(int slot) {
    return (slot == 11 || slot == 13 || slot == 7 || slot == 9 || slot == 1) ? 1
         : ((slot == 12 || slot == 14 || slot == 8 || slot == 10 || slot == 2) ? -1 : 0);
}
__global__ void k_flag_raise(unsigned int *sig_up, unsigned int *sig_dn, unsigned int epoch) {
    __threadfence_system();
    *(volatile unsigned int *)sig_up = epoch;
    *(volatile unsigned int *)sig_dn = epoch;
    __threadfence_system();
}

// ---- reductions ------------------------------------------------------------------------------------
// avedensity (collision.f90:497-498): per-block partial (count, sum) over fluid nodes, fixed order.
__global__ void __launch_bounds__(256) k_rho_partial(int lx, int xp, long long nrows, const double *rho,
                                                     const int32_t *solid, double *psum, long long *pcnt) {
    __shared__ double ss[256];
    __shared__ long long sc[256];
    double s = 0.0;
    long long c = 0;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x)
        for (int x = threadIdx.x; x < lx; x += 256) {
            const long long m = row * xp + x;
            if (!solid || solid[m] < 0) { s += rho[m]; c += 1; }
        }
    ss[threadIdx.x] = s; sc[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { ss[threadIdx.x] += ss[threadIdx.x + o]; sc[threadIdx.x] += sc[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { psum[blockIdx.x] = ss[0]; pcnt[blockIdx.x] = sc[0]; }
}
__global__ void k_rho_final(int nblk, const double *psum, const long long *pcnt, double *out_sum, long long *out_cnt) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0; long long c = 0;
        for (int i = 0; i < nblk; ++i) { s += psum[i]; c += pcnt[i]; }
        *out_sum = s; *out_cnt = c;
    }
}
__global__ void __launch_bounds__(BLOCK_X) k_rho_shift(int lx, int xp, double *rho, const double *mean_dev) {
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= lx) return;
    const long long row = blockIdx.y + (long long)gridDim.y * blockIdx.z;
    rho[row * xp + x] -= *mean_dev;            // collision.f90:508, every node
}

// statistc / statistc2 (saveload.f90:1241-1300, :1393-1421): per-x sums over (y,z) of 11 quantities and,
// as the 12th, the number of fluid nodes of the plane (statistc2's nfluid0).  Stage 1: each
// block owns BLOCK_X x-columns and a contiguous chunk of (y,z) rows; stage 2 adds chunks in order.
constexpr int NPROF = 12;
template <int RK>
__global__ void __launch_bounds__(BLOCK_X) k_profiles(Geom g, const double *A, double Fx, double Fy, double Fz,
                                                      const double *ffx, const double *ffy, const double *ffz,
                                                      const int32_t *solid, int rows_per_chunk, double *partial) {
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= g.lx) return;
    const long long nrows = (long long)g.ly * g.lz;
    const long long r0 = (long long)blockIdx.y * rows_per_chunk;
    const long long r1 = r0 + rows_per_chunk < nrows ? r0 + rows_per_chunk : nrows;
    double acc[NPROF];
#pragma unroll
    for (int q = 0; q < NPROF; ++q) acc[q] = 0.0;
    for (long long row = r0; row < r1; ++row) {
        const int y = (int)(row % g.ly), z = (int)(row / g.ly);
        if (solid && solid[row * g.xp + x] > 0) continue;
        const NodeIdx<unsigned long long> k = make_node<unsigned long long>(g, x, y, z + 1);
        double f[NPOP], r, a, b, c;
        gather19<RK>(A, g, k, f);
        if (ffx) { const long long m = row * g.xp + x; Fx = ffx[m]; Fy = ffy[m]; Fz = ffz[m]; }
        moments_strict(f, Fx, Fy, Fz, r, a, b, c);
        acc[0] += a; acc[1] += b; acc[2] += c;
        acc[3] += a * a; acc[4] += b * b; acc[5] += c * c;
        acc[6] += a * b; acc[7] += a * c; acc[8] += b * c;
        acc[9] += r; acc[10] += r * r;
        acc[11] += 1.0;
    }
#pragma unroll
    for (int q = 0; q < NPROF; ++q) partial[((long long)blockIdx.y * NPROF + q) * g.lx + x] = acc[q];
}
__global__ void __launch_bounds__(BLOCK_X) k_profiles_final(int lx, int nchunks, const double *partial, double *out) {
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= lx) return;
    const int q = blockIdx.y;
    double s = 0.0;
    for (int cidx = 0; cidx < nchunks; ++cidx) s += partial[((long long)cidx * NPROF + q) * lx + x];
    out[(long long)q * lx + x] = s;
}

// diag (saveload.f90:1507-1676): over the fluid nodes, count, sums of u and u^2, max |u| with the
// location of its first occurrence in the reference's loop order (z, y, x), max / min of rho.
// Stage 1 like k_profiles (a block owns BLOCK_X x-columns and a chunk of (y,z) rows), stage 2 one thread.
constexpr int NDIAG = 12;   // cnt, su, sv, sw, suu, svv, sww, vmax, vidx, rhomax, rhomin, (pad)
template <int RK>
__global__ void __launch_bounds__(BLOCK_X) k_diag(Geom g, const double *A, double Fx, double Fy, double Fz,
                                                  const double *ffx, const double *ffy, const double *ffz,
                                                  const int32_t *solid, int rows_per_chunk, double *partial) {
    __shared__ double sh[BLOCK_X][NDIAG];
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    const long long nrows = (long long)g.ly * g.lz;
    const long long r0 = (long long)blockIdx.y * rows_per_chunk;
    const long long r1 = r0 + rows_per_chunk < nrows ? r0 + rows_per_chunk : nrows;
    double acc[NDIAG];
#pragma unroll
    for (int q = 0; q < NDIAG; ++q) acc[q] = 0.0;
    acc[8] = -1.0; acc[9] = -1.0e300; acc[10] = 1.0e300;
    if (x < g.lx) {
        for (long long row = r0; row < r1; ++row) {
            const int y = (int)(row % g.ly), z = (int)(row / g.ly);
            if (solid && solid[row * g.xp + x] > 0) continue;
            const NodeIdx<unsigned long long> k = make_node<unsigned long long>(g, x, y, z + 1);
            double f[NPOP], r, a, b, c;
            gather19<RK>(A, g, k, f);
            if (ffx) { const long long m = row * g.xp + x; Fx = ffx[m]; Fy = ffy[m]; Fz = ffz[m]; }
            moments_strict(f, Fx, Fy, Fz, r, a, b, c);
            acc[0] += 1.0; acc[1] += a; acc[2] += b; acc[3] += c;
            acc[4] += a * a; acc[5] += b * b; acc[6] += c * c;
            const double vel = sqrt(a * a + b * b + c * c);
            if (vel > acc[7]) { acc[7] = vel; acc[8] = (double)(row * g.lx + x); }     // loop-order index, exact below 2^53
            acc[9] = r > acc[9] ? r : acc[9];
            acc[10] = r < acc[10] ? r : acc[10];
        }
    }
#pragma unroll
    for (int q = 0; q < NDIAG; ++q) sh[threadIdx.x][q] = acc[q];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < BLOCK_X; ++t) {
            for (int q = 0; q < 7; ++q) acc[q] += sh[t][q];
            if (sh[t][7] > acc[7] || (sh[t][7] == acc[7] && sh[t][8] >= 0.0 && (acc[8] < 0.0 || sh[t][8] < acc[8]))) {
                acc[7] = sh[t][7]; acc[8] = sh[t][8];
            }
            acc[9] = sh[t][9] > acc[9] ? sh[t][9] : acc[9];
            acc[10] = sh[t][10] < acc[10] ? sh[t][10] : acc[10];
        }
        double *o = partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * NDIAG;
        for (int q = 0; q < NDIAG; ++q) o[q] = acc[q];
    }
}
__global__ void k_diag_final(int npartial, const double *partial, double *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double acc[NDIAG];
    for (int q = 0; q < NDIAG; ++q) acc[q] = partial[q];
    for (int p = 1; p < npartial; ++p) {
        const double *s = partial + (long long)p * NDIAG;
        for (int q = 0; q < 7; ++q) acc[q] += s[q];
        if (s[7] > acc[7] || (s[7] == acc[7] && s[8] >= 0.0 && (acc[8] < 0.0 || s[8] < acc[8]))) { acc[7] = s[7]; acc[8] = s[8]; }
        acc[9] = s[9] > acc[9] ? s[9] : acc[9];
        acc[10] = s[10] < acc[10] ? s[10] : acc[10];
    }
    for (int q = 0; q < NDIAG; ++q) out[q] = acc[q];
}

}  // namespace d3q
