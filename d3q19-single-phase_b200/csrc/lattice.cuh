// lattice.cuh -- D3Q19 lattice tables in the reference's ordering (para.f90:178-206) and a
// compile-time loop so every direction index is a constant in the unrolled kernels.
#pragma once
#include <utility>

#if defined(__CUDACC__)
#define D3Q_HD __host__ __device__ __forceinline__
#else
#define D3Q_HD inline
#endif

namespace d3q {

constexpr int NPOP = 19;

D3Q_HD constexpr int dir_cx(int i) {
    constexpr int t[NPOP] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
    return t[i];
}
D3Q_HD constexpr int dir_cy(int i) {
    constexpr int t[NPOP] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
    return t[i];
}
D3Q_HD constexpr int dir_cz(int i) {
    constexpr int t[NPOP] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
    return t[i];
}
// ipopp, para.f90:206
D3Q_HD constexpr int dir_opp(int i) {
    constexpr int t[NPOP] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
    return t[i];
}

// The five populations that cross a z face (collision.f90:337-347).
D3Q_HD constexpr int face_pz(int s) { constexpr int t[5] = {5, 11, 12, 15, 16}; return t[s]; }
D3Q_HD constexpr int face_mz(int s) { constexpr int t[5] = {6, 13, 14, 17, 18}; return t[s]; }

template <class F, int... I>
D3Q_HD void static_for_impl(F &&f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
D3Q_HD void static_for(F &&f) {
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

// Lattice tables for kernels that index directions at run time (one link = one thread, every lane another
// direction): two bits per direction packed into a 64-bit literal and decoded in registers.  (A __constant__ array
// indexed by a per-lane direction serialises the constant cache: 18-way in k_beads_ibb, 150 us for 10^6 links.)
D3Q_HD constexpr unsigned long long pack_dirs(int which) {
    unsigned long long v = 0;
    for (int i = 0; i < NPOP; ++i) {
        const int c = which == 0 ? dir_cx(i) : (which == 1 ? dir_cy(i) : dir_cz(i));
        v |= (unsigned long long)(c + 1) << (2 * i);
    }
    return v;
}
D3Q_HD int rt_cx(int i) { return (int)((pack_dirs(0) >> (2 * i)) & 3ull) - 1; }
D3Q_HD int rt_cy(int i) { return (int)((pack_dirs(1) >> (2 * i)) & 3ull) - 1; }
D3Q_HD int rt_cz(int i) { return (int)((pack_dirs(2) >> (2 * i)) & 3ull) - 1; }
// ipopp (para.f90:204-206): (1,2)(3,4)(5,6) swap inside the pair, 7..10 -> 17-i, 11..14 -> 25-i, 15..18 -> 33-i
D3Q_HD int rt_opp(int i) {
    return i == 0 ? 0 : (i <= 6 ? ((i & 1) ? i + 1 : i - 1) : (i <= 10 ? 17 - i : (i <= 14 ? 25 - i : 33 - i)));
}

// MRT constants the path reads (para.f90:106-143); the fixed transform constants
// (coef*, val*, para.f90:143-170) are literals in collide.cuh.
struct Mrt {
    double s1, s2, s4, s9, s10, s13, s16;
    double omegepsl, omegepslj, omegxx;
};

}  // namespace d3q
