// particles.cuh -- device kernels of the particle path: solid mask and boundary links
// (beads_links), interpolated bounce-back with momentum-exchange force (beads_collision),
// short-range repulsion (beads_lubforce), rigid-body update (beads_move) and refill of the nodes
// a particle uncovered (beads_filling).  The names are the reference's timer names
// (var_inc.f90:166-168); the reference snapshot does NOT contain the code behind them
// (partlib.f90 is absent), so these kernels implement the published algorithms
// the reference's data structures point to -- Bouzidi et al. 2001 / Lallemand & Luo 2003 linear
// interpolated bounce-back, Galilean-invariant momentum exchange (Wen et al. 2014, Peng et al.
// 2016), equilibrium + non-equilibrium refill (Caiazzo 2008), Feng & Michaelides 2005 repulsion --
// and are checked against oracle/particles_oracle.c ("parity unpinned" against the reference).
//
// Geometry conventions are the reference's: node (ix,iy,iz) sits at (ix-0.5, iy-0.5, iz-0.5)
// (collision.f90:424-426); a node is solid iff |r - r_c| < rad; periodic images in y and z only
// (collision.f90:433-440); ibnodes = -1 fluid / > 0 solid, isnodes = owning particle id.  Here one
// ghosted int32 array `own[zg][y][x]` holds the owner id (1-based) or -1: own + plane is both the
// `solid` and the `isnodes` array the fluid kernels take.  The mask is a pure function of the
// replicated particle table, so every GPU fills its own ghost planes without any exchange.
//
// Integer work (mask, link list) uses non-contracted arithmetic (the R type of collide.cuh) so
// that it is bit-identical to the CPU checker.
#pragma once
#include "kernels.cuh"

namespace d3q {

struct PartGeom {
    Geom g;
    int nx, ny, nz, globalz;
    double rad;
};

__device__ __forceinline__ int wrap1(int j, int n) {
    int r = (j - 1) % n;
    if (r < 0) r += n;
    return r + 1;
}

struct BBox { int lo[3], n[3]; };

__device__ __forceinline__ BBox part_bbox(const PartGeom &pg, const double *c) {
    BBox b;
    int hi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        b.lo[d] = (int)floor((R(c[d]) - R(pg.rad) + R(0.5)).v) - 1;
        hi[d] = (int)ceil((R(c[d]) + R(pg.rad) + R(0.5)).v) + 1;
    }
    if (b.lo[0] < 1) b.lo[0] = 1;
    if (hi[0] > pg.nx) hi[0] = pg.nx;
#pragma unroll
    for (int d = 0; d < 3; ++d) b.n[d] = hi[d] - b.lo[d] + 1 > 0 ? hi[d] - b.lo[d] + 1 : 0;
    return b;
}

__device__ __forceinline__ double dist2_node(const double *c, int jx, int jy, int jz) {
    const R dx = R((double)jx - 0.5) - R(c[0]), dy = R((double)jy - 0.5) - R(c[1]), dz = R((double)jz - 0.5) - R(c[2]);
    return (dx * dx + dy * dy + dz * dz).v;
}

// local ghosted plane(s) of global plane iz (1..nz): up to two (a real plane and a periodic ghost copy)
__device__ __forceinline__ int local_planes(const PartGeom &pg, int iz, int out[3]) {
    int n = 0;
#pragma unroll
    for (int s = -1; s <= 1; ++s) {
        const int zg = iz - pg.globalz + s * pg.nz;
        if (zg >= 0 && zg <= pg.g.lz + 1) out[n++] = zg;
    }
    return n;
}

// The particle table is replicated on every GPU, the sweeps are launched for all particles: a particle whose bounding box
// (with its margin) misses this slab and its two ghost planes, in every periodic image, is dropped before anything else is
// computed -- on 8 slabs seven of eight particles are somebody else's (measured before this test existed: the
// particle-laden step grew from 1.96 ms on 1 GPU to 3.16 ms on 8 with 100 spheres per slab, profiles/r02g_eight_gpus.md).
__device__ __forceinline__ bool box_misses_slab(const PartGeom &pg, double cz) {
    const double lo = (double)pg.globalz - 1.0, hi = (double)(pg.globalz + pg.g.lz) + 1.0;     // node centres iz - 0.5, padded
    const double h = pg.rad + 3.5;                                                              // box half width and slack
#pragma unroll
    for (int s = -1; s <= 1; ++s) {
        const double z = cz + (double)s * (double)pg.nz;
        if (z + h >= lo && z - h <= hi) return false;
    }
    return true;
}

// ---- sweeps over a particle's bounding box: one WARP per box row -------------------------------------
// grid (npart, ceil(max rows / PART_WARPS)), PART_WARPS warps per block; warp w of block (p, by) owns row
// by * PART_WARPS + w of particle p's box (y fastest), its lanes walk along x: the mask is read and written in
// coalesced row pieces and a row that cannot hold anything is dropped before any memory access.
// (Round 1 gave each THREAD a contiguous chunk of the box: 19 scattered mask loads per node, two passes plus a scan
//  for the canonical order -- 536 us per step for 100 spheres of radius 15, profiles/r02a_launches_particles_summary.md.)
constexpr int PART_WARPS = 8;
inline int part_max_rows(double rad) { const int n = (int)ceil(2.0 * rad) + 6; return n * n; }      // box edge < 2 rad + 6

__device__ __forceinline__ bool box_row(const BBox &b, int &jy, int &jz) {
    const int row = (int)blockIdx.y * PART_WARPS + (int)(threadIdx.x >> 5);
    if (row >= b.n[1] * b.n[2]) return false;
    const int rz = row / b.n[1];
    jz = b.lo[2] + rz;
    jy = b.lo[1] + (row - rz * b.n[1]);
    return true;
}

// Is node (jx, iy, iz) -- iy, iz global and wrapped -- inside the sphere at c?  Evaluated with the unwrapped
// coordinates the sweep over c's own box uses, so that it is the very arithmetic of k_beads_cover (a box is
// smaller than the period: d3q19_particles_init checks)
__device__ __forceinline__ bool in_sphere_of(const PartGeom &pg, const double *c, const BBox &b, int jx, int iy, int iz, double r2) {
    int ry = (iy - b.lo[1]) % pg.ny, rz = (iz - b.lo[2]) % pg.nz;
    if (ry < 0) ry += pg.ny;
    if (rz < 0) rz += pg.nz;
    if (ry >= b.n[1] || rz >= b.n[2] || jx < b.lo[0] || jx >= b.lo[0] + b.n[0]) return false;
    return dist2_node(c, jx, b.lo[1] + ry, b.lo[2] + rz) < r2;
}

// ---- beads_links, part 1: the solid mask follows the particles ---------------------------------------
// ONE ghosted owner array, updated in place by two sweeps (no second array, no clearing of 4 B per node and step):
//   k_beads_uncover  over the box of the position the mask was built with (ypmask): a node owned by p that the
//                    NEW sphere of p no longer covers becomes -(p+2) -- "uncovered by p in this update", fluid to
//                    every reader (they test own > 0) -- and goes onto the refill list; markers left by the previous
//                    update (they lie within rad + 1 of ypmask) go back to -1;
//   k_beads_cover    over the box of the new position: nodes inside the sphere get p+1, the lowest id wins where
//                    spheres overlap (also over a marker: another particle may cover what p left).
// After both, own == -1: fluid before and after; own == -(q+2): solid before, fluid now (what beads_filling
// rebuilds, and never a refill source); own > 0: solid now.  This is what the CPU checker's pair of arrays (mask before / after the move) encodes.
struct FillList {
    uint32_t *node;               // ghosted in-slab index of an uncovered node of this slab
    int32_t *part;                // the particle that left it, 1-based
    unsigned long long *count;
    long long cap;
};

__global__ void __launch_bounds__(32 * PART_WARPS) k_beads_uncover(PartGeom pg, int npart, const double *ypmask, const double *ypglb,
                                                                   int32_t *own, FillList F) {
    const int p = blockIdx.x, lane = threadIdx.x & 31;
    const double *c0 = ypmask + 3 * p, *c1 = ypglb + 3 * p;
    if (box_misses_slab(pg, c0[2])) return;
    const BBox b = part_bbox(pg, c0), b1 = part_bbox(pg, c1);
    int jy, jz;
    if (!box_row(b, jy, jz)) return;
    const double r2 = (R(pg.rad) * R(pg.rad)).v;
    {   // nothing of p and none of its markers lies further than rad + 1 from c0 (a particle moves less than a node per step)
        const double dy = ((double)jy - 0.5) - c0[1], dz = ((double)jz - 0.5) - c0[2], rr = pg.rad + 1.5;
        if (dy * dy + dz * dz > rr * rr) return;
    }
    const int iy = wrap1(jy, pg.ny), iz = wrap1(jz, pg.nz);
    int zs[3];
    const int nzs = local_planes(pg, iz, zs);
    const int32_t marker = -(p + 2);
    for (int rx = lane; rx < b.n[0]; rx += 32) {
        const int jx = b.lo[0] + rx;
        for (int q = 0; q < nzs; ++q) {
            const long long n = (long long)(jx - 1) + (long long)pg.g.xp * ((iy - 1) + (long long)pg.g.ly * zs[q]);
            const int32_t v = own[n];
            if (v == marker) {
                own[n] = -1;
            } else if (v == p + 1 && !in_sphere_of(pg, c1, b1, jx, iy, iz, r2)) {
                own[n] = marker;
                if (zs[q] >= 1 && zs[q] <= pg.g.lz) {
                    const unsigned long long w = atomicAdd(F.count, 1ull);
                    if ((long long)w < F.cap) { F.node[w] = (uint32_t)n; F.part[w] = p + 1; }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(32 * PART_WARPS) k_beads_cover(PartGeom pg, int npart, const double *ypglb, int32_t *own) {
    const int p = blockIdx.x, lane = threadIdx.x & 31;
    const double *c = ypglb + 3 * p;
    if (box_misses_slab(pg, c[2])) return;
    const BBox b = part_bbox(pg, c);
    int jy, jz;
    if (!box_row(b, jy, jz)) return;
    const double r2 = (R(pg.rad) * R(pg.rad)).v;
    {
        const double dy = ((double)jy - 0.5) - c[1], dz = ((double)jz - 0.5) - c[2], rr = pg.rad + 0.01;      // conservative
        if (dy * dy + dz * dz > rr * rr) return;
    }
    const int iy = wrap1(jy, pg.ny), iz = wrap1(jz, pg.nz);
    int zs[3];
    const int nzs = local_planes(pg, iz, zs);
    for (int rx = lane; rx < b.n[0]; rx += 32) {
        const int jx = b.lo[0] + rx;
        if (!(dist2_node(c, jx, jy, jz) < r2)) continue;
        for (int q = 0; q < nzs; ++q) {
            int32_t *a = own + ((long long)(jx - 1) + (long long)pg.g.xp * ((iy - 1) + (long long)pg.g.ly * zs[q]));
            int32_t old = *a;                       // the lowest id wins where particles overlap
            while (old < 0 || old > p + 1) {
                const int32_t assumed = old;
                old = atomicCAS(a, assumed, p + 1);
                if (old == assumed) break;
            }
        }
    }
}

// ---- beads_links, part 2: boundary links ------------------------------------------------------------
// A link = (fluid node of this slab, direction whose neighbour is owned by particle p).  Every particle has its own
// SEGMENT of `cap` entries in the list and its own counter: warps append their row's links with one atomic per 32
// nodes, so inside a segment the order between rows depends on the run (inside a pass it is x, then direction) -- the
// list is a set.  Consumers do not depend on the order (k_beads_ibb: one thread per link, every link writes its own
// slot; all links of a block belong to one particle, so the force is reduced per warp before it is added);
// d3q19_get_links hands the segments out concatenated and tests compare after a canonical sort (SURVEY.md appendix B:
// "bit-exact after a canonical sort").  The fraction q of a link is NOT stored: it is one square root and one division,
// cheaper to recompute per link by whoever needs it (without divergence) than to evaluate inside the sweep, where a
// warp would run all 18 directions' roots for the union of its lanes' links (measured: 256 us of a 2.7 ms step).
struct Links {
    uint32_t *node;            // ghosted in-slab index of the fluid node, [npart][cap]
    int32_t *dir;              // direction pointing into the solid
    unsigned long long *count; // links of particle p on this slab (may exceed cap: the excess was dropped -> error)
    long long cap;             // entries per particle
};

// fraction of the link from (jx,jy,jz) along IP that lies in the fluid: smallest root of
// |x_f + t c - r_c|^2 = rad^2, non-contracted arithmetic (bit-identical to the CPU checker)
template <int IP>
__device__ __forceinline__ double link_q(const double *c, int jx, int jy, int jz, double r2) {
    constexpr int cx = dir_cx(IP), cy = dir_cy(IP), cz = dir_cz(IP);
    const R dx = R((double)jx - 0.5) - R(c[0]), dy = R((double)jy - 0.5) - R(c[1]), dz = R((double)jz - 0.5) - R(c[2]);
    const R a((double)(cx * cx + cy * cy + cz * cz));
    const R b = R(2.0) * (R((double)cx) * dx + R((double)cy) * dy + R((double)cz) * dz);
    const R cc = dx * dx + dy * dy + dz * dz - R(r2);
    double disc = (b * b - R(4.0) * a * cc).v;
    if (disc < 0.0) disc = 0.0;
    double q = __ddiv_rn((-b - R(__dsqrt_rn(disc))).v, (R(2.0) * a).v);
    if (q < 0.0) q = 0.0;
    if (q > 1.0) q = 1.0;
    return q;
}

// the same for a run-time direction (one thread per link) and the coordinates of a link node as the sweep over the
// particle's box saw them (jy, jz unwrapped next to the centre): what k_beads_ibb and the export use
__device__ __forceinline__ double link_q_rt(const double *c, int jx, int jy, int jz, double r2, int ip) {
    const int cx = rt_cx(ip), cy = rt_cy(ip), cz = rt_cz(ip);
    const R dx = R((double)jx - 0.5) - R(c[0]), dy = R((double)jy - 0.5) - R(c[1]), dz = R((double)jz - 0.5) - R(c[2]);
    const R a((double)(cx * cx + cy * cy + cz * cz));
    const R b = R(2.0) * (R((double)cx) * dx + R((double)cy) * dy + R((double)cz) * dz);
    const R cc = dx * dx + dy * dy + dz * dz - R(r2);
    double disc = (b * b - R(4.0) * a * cc).v;
    if (disc < 0.0) disc = 0.0;
    double q = __ddiv_rn((-b - R(__dsqrt_rn(disc))).v, (R(2.0) * a).v);
    if (q < 0.0) q = 0.0;
    if (q > 1.0) q = 1.0;
    return q;
}
// (jx, jy, jz) of the node with ghosted in-slab index n, unwrapped into particle p's box
__device__ __forceinline__ void link_node_coords(const PartGeom &pg, const double *c, uint32_t n, int &jx, int &jy, int &jz) {
    const Geom &g = pg.g;
    const int x = (int)(n % (uint32_t)g.xp), y = (int)((n / (uint32_t)g.xp) % (uint32_t)g.ly), zg = (int)(n / ((uint32_t)g.xp * (uint32_t)g.ly));
    const BBox b = part_bbox(pg, c);
    int ry = (y + 1 - b.lo[1]) % pg.ny, rz = (zg + pg.globalz - b.lo[2]) % pg.nz;
    if (ry < 0) ry += pg.ny;
    if (rz < 0) rz += pg.nz;
    jx = x + 1; jy = b.lo[1] + ry; jz = b.lo[2] + rz;
}

// One warp per box row.  The link nodes of a row sit in two short runs of columns where the row enters and leaves the
// shell rad <= d < rad + 1.5 (one run where it only grazes it); each half of the warp takes a 16-column window over one
// run -- 14 nodes and their two x-neighbours -- so that one pass usually covers the row.  The nine mask rows around the row
// (y-1..y+1, z-1..z+1) are read as coalesced pieces and the 18 neighbour owners of a node come from its own registers
// (c_x = 0) or from the adjacent lane (shuffle); list space is reserved with one atomic per pass.  The kernel is bound by
// instruction issue (ncu, r02f: 105 M warp instructions, 74 % issue slots busy, 130 us for 100 spheres of radius 15 with
// 30-column pieces, two per row, and a direction-major order that cost 36 votes per row and bought k_beads_ibb nothing).
__device__ __forceinline__ int wrap_near(int j, int n) { return j < 1 ? j + n : (j > n ? j - n : j); }     // |j - [1,n]| < n

__global__ void __launch_bounds__(32 * PART_WARPS) k_beads_links(PartGeom pg, int npart, const double *ypglb, const int32_t *own, Links L) {
    const int p = blockIdx.x, lane = threadIdx.x & 31;
    const double *c = ypglb + 3 * p;
    if (box_misses_slab(pg, c[2])) return;
    const BBox b = part_bbox(pg, c);
    int jy, jz;
    if (!box_row(b, jy, jz)) return;
    const double r2 = (R(pg.rad) * R(pg.rad)).v;
    // Only a thin shell can hold link nodes: a node with d < rad was claimed by k_beads_cover (same arithmetic) for this
    // or a lower-numbered particle, and a neighbour owned by this particle lies within rad, so the node itself within
    // rad + |c_i| <= rad + sqrt(2) < rad + 1.5.  Rows outside the shell's (y,z) shadow are dropped here.
    const double rshell2 = (pg.rad + 1.5) * (pg.rad + 1.5);
    const double dy = ((double)jy - 0.5) - c[1], dz = ((double)jz - 0.5) - c[2];
    const double rho2 = dy * dy + dz * dz;
    if (rho2 > rshell2) return;
    const int iy = wrap_near(jy, pg.ny), iz = wrap_near(jz, pg.nz);        // a box is smaller than the period
    const int zg = iz - pg.globalz;                  // links belong to the GPU that owns the fluid node
    if (zg < 1 || zg > pg.g.lz) return;
    // columns that can hold a candidate: |x - 0.5 - c_x| in [a_in, a_out); one column of slack on either side of a run,
    // the exact test below stays the authority
    const double a_out = sqrt(rshell2 - rho2), a_in = rho2 < r2 ? sqrt(r2 - rho2) : 0.0;
    int xa0 = (int)floor(c[0] + 0.5 - a_out) - 1, xa1 = (int)ceil(c[0] + 0.5 - a_in) + 1;      // where the row enters
    int xb0 = (int)floor(c[0] + 0.5 + a_in) - 1, xb1 = (int)ceil(c[0] + 0.5 + a_out) + 1;      // where it leaves
    if (xb0 <= xa1 + 1) { xa1 = xb1; xb0 = 1; xb1 = 0; }                     // one run (the row grazes the shell)
    if (xa0 < b.lo[0]) xa0 = b.lo[0];
    if (xb1 > b.lo[0] + b.n[0] - 1) xb1 = b.lo[0] + b.n[0] - 1;
    if (xa1 > b.lo[0] + b.n[0] - 1) xa1 = b.lo[0] + b.n[0] - 1;
    // the nine rows: y-1, y, y+1 (periodic) x z-1, z, z+1 (ghost planes carry the mask too); 32-bit: d3q19_particles_init
    // refuses slabs whose populations need 64-bit indices
    uint32_t rowbase[3][3];
    {
        const int ky[3] = {wrap_near(iy - 1, pg.ny), iy, wrap_near(iy + 1, pg.ny)};
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int bz = 0; bz < 3; ++bz) rowbase[a][bz] = (uint32_t)pg.g.xp * (uint32_t)((ky[a] - 1) + pg.g.ly * (zg + bz - 1));
    }
    const unsigned full = 0xffffffffu;
    const int half = lane >> 4, hl = lane & 15;      // lanes 0-15: the first run, 16-31: the second; hl 0 and 15: x-neighbours
    const long long seg = (long long)p * L.cap;
    for (int k = 0; xa0 + 14 * k <= xa1 || xb0 + 14 * k <= xb1; ++k) {       // warp-uniform trip count (the shuffles need every lane)
        const int x0 = (half ? xb0 : xa0) + 14 * k, x1 = half ? xb1 : xa1;
        const int jx = x0 + hl - 1;
        const bool inside = jx >= 1 && jx <= pg.nx;
        bool cand = false;
        if (hl >= 1 && hl <= 14 && jx <= x1 && inside) {
            const double d2 = dist2_node(c, jx, jy, jz);
            cand = !(d2 < r2) && d2 < rshell2;
        }
        if (!__any_sync(full, cand)) continue;
        int32_t o[3][3];                             // owners of the nine rows at column jx (-2 beyond a channel wall)
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int bz = 0; bz < 3; ++bz) o[a][bz] = inside ? own[rowbase[a][bz] + (uint32_t)(jx - 1)] : -2;
        if (o[1][1] > 0) cand = false;               // the node itself is solid (this or another particle)
        unsigned bits = 0u;                           // bit ip-1: the neighbour along ip is owned by p
        static_for<NPOP - 1>([&](auto ic) {
            constexpr int ip = decltype(ic)::value + 1;
            constexpr int cx = dir_cx(ip), cy = dir_cy(ip), cz = dir_cz(ip);
            int32_t v = o[1 + cy][1 + cz];
            if (cx != 0) v = __shfl_sync(full, v, (lane + cx) & 31);         // (never across the halves: hl 0 and 15 are no candidates)
            if (cand && v == p + 1) bits |= 1u << (ip - 1);
        });
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int of = 1; of < 32; of <<= 1) {
            const int t = __shfl_up_sync(full, incl, of);
            if (lane >= of) incl += t;
        }
        const int total = __shfl_sync(full, incl, 31);
        if (total == 0) continue;
        unsigned long long base = 0;
        if (lane == 31) base = atomicAdd(L.count + p, (unsigned long long)total);
        base = __shfl_sync(full, base, 31);
        long long w = (long long)base + incl - cnt;
        const uint32_t n = rowbase[1][1] + (uint32_t)(jx - 1);
        for (unsigned rest = bits; rest; rest &= rest - 1) {               // the set bits, lowest direction first
            if (w < L.cap) { L.node[seg + w] = n; L.dir[seg + w] = __ffs(rest); }
            ++w;
        }
    }
}

// ---- beads_collision: interpolated bounce-back + momentum exchange ---------------------------------
// One thread per link, after the step (and after its halo exchange).  Where the three post-collision
// values live and where the result goes, by storage state:
//   READ_PULL_NAT  (AB):        f*_i(x) = S[i][x];        result -> S[opp i][x_s]   (x_f pulls it from there)
//   READ_PULL_SWAP (AA, even):  f*_i(x) = S[opp i][x];    result -> S[i][x_s]
//   READ_DIRECT    (AA, odd):   f*_i(x) = S[i][x + c_i];  result -> S[opp i][x_f]   (already streamed)
struct IbbParams {
    PartGeom pg;
    double *S;
    const int32_t *own;
    Links L;                  // the counts stay on the device: no host round trip between links and IBB
    const double *ypglb, *wp, *omgp;
    double rho0;
    double *fHIp, *torqp;     // (3,npart), accumulated with atomics
};

template <int RK>
__device__ __forceinline__ long long post_addr(const Geom &g, int i, long long n, long long nplus) {
    // address of f*_i of the node with in-slab index n; nplus = index of that node + c_i
    if (RK == READ_PULL_NAT) return (long long)i * g.slab + n;
    if (RK == READ_PULL_SWAP) return (long long)rt_opp(i) * g.slab + n;
    return (long long)i * g.slab + nplus;
}

template <int RK>
__global__ void __launch_bounds__(128) k_beads_ibb(const __grid_constant__ IbbParams P) {
    const Geom &g = P.pg.g;
    // grid (blocks of a segment, npart): all links of a block belong to particle blockIdx.y
    const int p = blockIdx.y;
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double F[6] = {0, 0, 0, 0, 0, 0};
    const long long nlink = (long long)P.L.count[p] < P.L.cap ? (long long)P.L.count[p] : P.L.cap;
    if ((long long)blockIdx.x * blockDim.x >= nlink) return;          // whole block beyond the segment's end
    if (l < nlink) {
        const uint32_t n = P.L.node[(long long)p * P.L.cap + l];
        const int ip = P.L.dir[(long long)p * P.L.cap + l], io = rt_opp(ip);
        double q;
        {
            int jx, jy, jz;
            link_node_coords(P.pg, P.ypglb + 3 * p, n, jx, jy, jz);
            q = link_q_rt(P.ypglb + 3 * p, jx, jy, jz, (R(P.pg.rad) * R(P.pg.rad)).v, ip);
        }
        const int x = (int)(n % (uint32_t)g.xp), y = (int)((n / (uint32_t)g.xp) % (uint32_t)g.ly),
                  zg = (int)(n / ((uint32_t)g.xp * (uint32_t)g.ly));
        const int cx = rt_cx(ip), cy = rt_cy(ip), cz = rt_cz(ip);
        const double ww = ip <= 6 ? 1.0 / 18.0 : 1.0 / 36.0;
        // neighbours along the link: x_s = x_f + c, x_b = x_f - c (periodic y; z through wrap or ghosts)
        auto wrapz = [&](int z) { return z < 1 ? g.zlo_src : (z > g.lz ? g.zhi_src : z); };
        auto wrapy = [&](int yy) { return yy < 0 ? g.ly - 1 : (yy >= g.ly ? 0 : yy); };
        const long long ns = (long long)(x + cx) + (long long)g.xp * (wrapy(y + cy) + (long long)g.ly * wrapz(zg + cz));
        const bool back_in = (x - cx) >= 0 && (x - cx) < g.lx;
        const long long nb = back_in ? (long long)(x - cx) + (long long)g.xp * (wrapy(y - cy) + (long long)g.ly * wrapz(zg - cz)) : (long long)n;
        // wall point relative to the particle centre (nearest image)
        double c0 = P.ypglb[3 * p], c1 = P.ypglb[3 * p + 1], c2 = P.ypglb[3 * p + 2];
        const double xf0 = (double)(x + 1) - 0.5, xf1 = (double)(y + 1) - 0.5, xf2 = (double)(zg + P.pg.globalz) - 0.5;
        if (c1 - xf1 > 0.5 * P.pg.ny) c1 -= P.pg.ny;
        if (c1 - xf1 < -0.5 * P.pg.ny) c1 += P.pg.ny;
        if (c2 - xf2 > 0.5 * P.pg.nz) c2 -= P.pg.nz;
        if (c2 - xf2 < -0.5 * P.pg.nz) c2 += P.pg.nz;
        const double rx = xf0 + q * cx - c0, ry = xf1 + q * cy - c1, rz = xf2 + q * cz - c2;
        const double uwx = P.wp[3 * p] + (P.omgp[3 * p + 1] * rz - P.omgp[3 * p + 2] * ry);
        const double uwy = P.wp[3 * p + 1] + (P.omgp[3 * p + 2] * rx - P.omgp[3 * p] * rz);
        const double uwz = P.wp[3 * p + 2] + (P.omgp[3 * p] * ry - P.omgp[3 * p + 1] * rx);
        const double delta = 6.0 * ww * P.rho0 * (-(cx * uwx + cy * uwy + cz * uwz));
        const double fs_i = P.S[post_addr<RK>(g, ip, n, ns)];                    // f*_i(x_f)
        double fnew;
        if (q >= 0.5) {
            // f*_opp(x_f); at a channel wall behind the node it has bounced back into direction ip
            double fs_o;
            if (back_in) fs_o = P.S[post_addr<RK>(g, io, n, nb)];
            else fs_o = (RK == READ_DIRECT) ? P.S[(long long)ip * g.slab + n] : P.S[post_addr<RK>(g, io, n, n)];
            const double i2q = 1.0 / (2.0 * q);
            fnew = i2q * fs_i + (2.0 * q - 1.0) * i2q * fs_o + i2q * delta;
        } else {
            const bool have_ff = back_in && P.own[nb] < 0;
            if (have_ff) fnew = 2.0 * q * fs_i + (1.0 - 2.0 * q) * P.S[post_addr<RK>(g, ip, nb, n)] + delta;   // f*_i(x_f - c_i)
            else fnew = fs_i + delta;
        }
        // where x_f will find f_opp(i) at the next step
        long long dst;
        if (RK == READ_PULL_NAT) dst = (long long)io * g.slab + ns;
        else if (RK == READ_PULL_SWAP) dst = (long long)ip * g.slab + ns;
        else dst = (long long)io * g.slab + n;
        P.S[dst] = fnew;
        const double fin = fs_i + ww * P.rho0, fout = fnew + ww * P.rho0;
        F[0] = (cx - uwx) * fin - (-cx - uwx) * fout;
        F[1] = (cy - uwy) * fin - (-cy - uwy) * fout;
        F[2] = (cz - uwz) * fin - (-cz - uwz) * fout;
        F[3] = ry * F[2] - rz * F[1];
        F[4] = rz * F[0] - rx * F[2];
        F[5] = rx * F[1] - ry * F[0];
    }
    // one particle per block: reduce over the warp, one atomic per component per warp
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double v = F[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
        F[k] = v;
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicAdd(P.fHIp + 3 * p + k, F[k]); atomicAdd(P.torqp + 3 * p + k, F[3 + k]); }
    }
}

// ---- beads_filling: nodes solid before the move and fluid after it ----------------------------------
struct FillParams {
    PartGeom pg;
    double *S;
    const int32_t *own;           // after the mask update: -1 fluid before and after, -(q+2) uncovered by q, > 0 solid
    FillList F;                   // the nodes k_beads_uncover found
    const double *ypglb, *wp, *omgp;
    unsigned long long *nfilled;
    // z-slab runs: the 19 canonical populations of the neighbours' planes next to the faces ([i][y][x], pitch xp),
    // lo = the lower neighbour's plane lz, hi = the upper neighbour's plane 1 (k_plane_gather + send/recv before the
    // refill).  A ghost plane of the population array holds only the 5 populations that cross the face, not a
    // node's 19; with these copies a refill next to a face sees the same source nodes as on a single domain.
    const double *ghost_lo, *ghost_hi;
};

// canonical populations of two planes -> out[i][y][x] (pitch xp), whatever the storage phase: blockIdx.z = 0 -> plane
// zg0 into out0, 1 -> plane zg1 into out1 (the two faces of a slab in one launch)
template <int RK>
__global__ void __launch_bounds__(BLOCK_X) k_plane_gather(Geom g, const double *A, double *out0, int zg0, double *out1, int zg1) {
    const int x = blockIdx.x * BLOCK_X + threadIdx.x;
    if (x >= g.lx) return;
    double *out = blockIdx.z == 0 ? out0 : out1;
    const NodeIdx<unsigned long long> k = make_node<unsigned long long>(g, x, blockIdx.y, blockIdx.z == 0 ? zg0 : zg1);
    double f[NPOP];
    gather19<RK>(A, g, k, f);
    const long long o = (long long)blockIdx.y * g.xp + x;
#pragma unroll
    for (int i = 0; i < NPOP; ++i) out[(long long)i * g.plane + o] = f[i];
}

__device__ __forceinline__ void feq19(double rho, double ux, double uy, double uz, double (&fe)[NPOP]) {
    const double usqr = 1.5 * (ux * ux + uy * uy + uz * uz);
    static_for<NPOP>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        if (i == 0) fe[0] = (1.0 / 3.0) * (rho - usqr);
        else {
            const double G = dir_cx(i) * ux + dir_cy(i) * uy + dir_cz(i) * uz;
            fe[i] = (i <= 6 ? 1.0 / 18.0 : 1.0 / 36.0) * (rho + 3.0 * G + 4.5 * G * G - usqr);
        }
    });
}

template <int RK>
__global__ void __launch_bounds__(128) k_beads_fill(const __grid_constant__ FillParams P) {
    const PartGeom &pg = P.pg;
    const Geom &g = pg.g;
    const long long nlist = (long long)*P.F.count < P.F.cap ? (long long)*P.F.count : P.F.cap;
    // the list length is known on the device only: a fixed grid strides over it
    for (long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x; l < nlist; l += (long long)gridDim.x * blockDim.x) {
        const uint32_t nn32 = P.F.node[l];
        const int p = P.F.part[l] - 1;
        const long long n = (long long)nn32;
        if (P.own[n] != -(p + 2)) continue;               // another particle covers the node now
        const int x = (int)(nn32 % (uint32_t)g.xp), y = (int)((nn32 / (uint32_t)g.xp) % (uint32_t)g.ly),
                  zg = (int)(nn32 / ((uint32_t)g.xp * (uint32_t)g.ly));
        const int jx = x + 1, iy = y + 1, iz = zg + pg.globalz;
        // surface velocity of the particle that uncovered the node, at the node (state after the move)
        double c0 = P.ypglb[3 * p], c1 = P.ypglb[3 * p + 1], c2 = P.ypglb[3 * p + 2];
        const double xf0 = (double)jx - 0.5, xf1 = (double)iy - 0.5, xf2 = (double)iz - 0.5;
        if (c1 - xf1 > 0.5 * pg.ny) c1 -= pg.ny;
        if (c1 - xf1 < -0.5 * pg.ny) c1 += pg.ny;
        if (c2 - xf2 > 0.5 * pg.nz) c2 -= pg.nz;
        if (c2 - xf2 < -0.5 * pg.nz) c2 += pg.nz;
        const double rx = xf0 - c0, ry = xf1 - c1, rz = xf2 - c2;
        const double uwx = P.wp[3 * p] + (P.omgp[3 * p + 1] * rz - P.omgp[3 * p + 2] * ry);
        const double uwy = P.wp[3 * p + 1] + (P.omgp[3 * p + 2] * rx - P.omgp[3 * p] * rz);
        const double uwz = P.wp[3 * p + 2] + (P.omgp[3 * p] * ry - P.omgp[3 * p + 1] * rx);
        double rsum = 0.0, best = -2.0;
        int nn = 0, jbest = 0;
        double fbest[NPOP];
        for (int ip = 1; ip < NPOP; ++ip) {
            const int cx = rt_cx(ip), cy = rt_cy(ip), cz = rt_cz(ip);
            const int kx = x + cx;
            if (kx < 0 || kx >= g.lx) continue;
            const int ky = (y + cy < 0) ? g.ly - 1 : (y + cy >= g.ly ? 0 : y + cy);
            int kz = zg + cz;
            const double *gh = nullptr;
            if (g.zlo_src != 0) kz = kz < 1 ? g.lz : (kz > g.lz ? 1 : kz);       // single slab: periodic wrap
            else if (kz < 1 || kz > g.lz) {                                        // a neighbour across a slab face
                gh = kz < 1 ? P.ghost_lo : P.ghost_hi;
                if (!gh) continue;
            }
            const long long m = (long long)kx + (long long)g.xp * (ky + (long long)g.ly * kz);     // the masks are ghosted
            if (P.own[m] != -1) continue;                 // solid now, or solid before this update (a marker)
            double fm[NPOP];
            if (gh) {
                const long long o = (long long)ky * g.xp + kx;
#pragma unroll
                for (int i = 0; i < NPOP; ++i) fm[i] = gh[(long long)i * g.plane + o];
            } else {
                const NodeIdx<unsigned long long> k = make_node<unsigned long long>(g, kx, ky, kz);
                gather19<RK>(P.S, g, k, fm);
            }
            double r = 0.0;
#pragma unroll
            for (int i = 0; i < NPOP; ++i) r += fm[i];
            rsum += r; ++nn;
            const double cn = (cx * rx + cy * ry + cz * rz) / sqrt((double)(cx * cx + cy * cy + cz * cz));
            if (cn > best) {
                best = cn; jbest = ip;
#pragma unroll
                for (int i = 0; i < NPOP; ++i) fbest[i] = fm[i];
            }
        }
        const double rbar = nn ? rsum / (double)nn : 0.0;
        double out[NPOP];
        feq19(rbar, uwx, uwy, uwz, out);
        if (jbest) {
            double r = 0.0, jx_ = 0.0, jy_ = 0.0, jz_ = 0.0, fe2[NPOP];
            static_for<NPOP>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                r += fbest[i]; jx_ += dir_cx(i) * fbest[i]; jy_ += dir_cy(i) * fbest[i]; jz_ += dir_cz(i) * fbest[i];
            });
            feq19(r, jx_, jy_, jz_, fe2);
#pragma unroll
            for (int i = 0; i < NPOP; ++i) out[i] += fbest[i] - fe2[i];
        }
        // write each population where this node will read it from
        const NodeIdx<unsigned long long> k = make_node<unsigned long long>(g, x, y, zg);
        static_for<NPOP>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            P.S[Gather<RK, i>::offset(g, k)] = out[i];
        });
        atomicAdd(P.nfilled, 1ull);
    }
}

// ---- beads_lubforce / beads_move: npart threads ------------------------------------------------------
struct LubParams { double mingap, mingap_w, stf0, stf1, stf0_w, stf1_w, fscale; };

// one WARP per particle i: the lanes share the partners j (a particle's partner loop is a chain of square roots), the three
// components are reduced over the warp, lane 0 adds the two walls and stores
__device__ __forceinline__ void lubforce_warp(const PartGeom &pg, int npart, const double *ypglb, const LubParams &lp, double *flubp, int i) {
    const int lane = threadIdx.x & 31;
    const double *a = ypglb + 3 * i;
    const double Rr = pg.rad;
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    for (int j = lane; j < npart; j += 32) {
        if (j == i) continue;
        const double *b = ypglb + 3 * j;
        double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        if (dy > 0.5 * pg.ny) dy -= pg.ny;
        if (dy < -0.5 * pg.ny) dy += pg.ny;
        if (dz > 0.5 * pg.nz) dz -= pg.nz;
        if (dz < -0.5 * pg.nz) dz += pg.nz;
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        const double gap = d - 2.0 * Rr;
        if (gap >= lp.mingap || d == 0.0) continue;
        double mag = lp.fscale / lp.stf0 * ((gap - lp.mingap) / lp.mingap) * ((gap - lp.mingap) / lp.mingap);
        if (gap < 0.0) mag += lp.fscale / lp.stf1 * (-gap / lp.mingap);
        f0 += mag * dx / d; f1 += mag * dy / d; f2 += mag * dz / d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        f0 += __shfl_xor_sync(0xffffffffu, f0, o);
        f1 += __shfl_xor_sync(0xffffffffu, f1, o);
        f2 += __shfl_xor_sync(0xffffffffu, f2, o);
    }
    if (lane == 0) {
        for (int s = 0; s < 2; ++s) {
            const double dxw = s == 0 ? a[0] : a[0] - (double)pg.nx;
            const double gap = fabs(dxw) - Rr;
            if (gap >= lp.mingap_w) continue;
            double mag = lp.fscale / lp.stf0_w * ((gap - lp.mingap_w) / lp.mingap_w) * ((gap - lp.mingap_w) / lp.mingap_w);
            if (gap < 0.0) mag += lp.fscale / lp.stf1_w * (-gap / lp.mingap_w);
            f0 += (dxw >= 0.0 ? mag : -mag);
        }
        flubp[3 * i] = f0; flubp[3 * i + 1] = f1; flubp[3 * i + 2] = f2;
    }
}

struct MoveParams {
    double amp, aip;
    double g0, g1, g2;
    const double *fHIp, *torqp, *flubp;
    double *forcepp, *torqpp, *ypglb, *wp, *omgp, *thetap;
};

__device__ __forceinline__ void move_one(const PartGeom &pg, const MoveParams &M, int p) {
    const double gf[3] = {M.g0, M.g1, M.g2};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int k = 3 * p + d;
        const double F = 0.5 * (M.fHIp[k] + M.forcepp[k]) + M.flubp[k] + gf[d];
        const double T = 0.5 * (M.torqp[k] + M.torqpp[k]);
        const double wnew = M.wp[k] + F / M.amp;
        const double onew = M.omgp[k] + T / M.aip;
        M.ypglb[k] += 0.5 * (M.wp[k] + wnew);
        M.thetap[k] += 0.5 * (M.omgp[k] + onew);
        M.wp[k] = wnew; M.omgp[k] = onew;
        M.forcepp[k] = M.fHIp[k]; M.torqpp[k] = M.torqp[k];
    }
    double *c = M.ypglb + 3 * p;
    if (c[1] >= (double)pg.ny) c[1] -= pg.ny;
    if (c[1] < 0.0) c[1] += pg.ny;
    if (c[2] >= (double)pg.nz) c[2] -= pg.nz;
    if (c[2] < 0.0) c[2] += pg.nz;
}

// beads_lubforce and beads_move of all particles in ONE block (npart is a few hundred): every repulsion force is formed
// from the positions before anybody moves, then the barrier, then the rigid-body update
__global__ void __launch_bounds__(1024) k_beads_lubmove(PartGeom pg, int npart, const double *ypglb, LubParams lp, double *flubp,
                                                        MoveParams M, int do_lub, int do_move) {
    // (do_lub without do_move may be launched on several blocks: the partner loop is O(npart^2))
    if (do_lub)
        for (int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < npart; i += (int)((gridDim.x * blockDim.x) >> 5))
            lubforce_warp(pg, npart, ypglb, lp, flubp, i);
    __syncthreads();
    if (do_move)
        for (int p = threadIdx.x; p < npart; p += blockDim.x) move_one(pg, M, p);
}

// link list -> host-friendly, segments concatenated: global 1-based node coordinates, direction, particle, q;
// grid (blocks of a segment, npart), offset[p] = where particle p's links start in the output
__global__ void k_links_export(PartGeom pg, Links L, const double *ypglb, const long long *offset, int32_t *ox, int32_t *oy, int32_t *oz,
                               int32_t *odir, int32_t *opart, double *oq) {
    const Geom &g = pg.g;
    const int p = blockIdx.y;
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nlink = (long long)L.count[p] < L.cap ? (long long)L.count[p] : L.cap;
    if (l >= nlink) return;
    const uint32_t n = L.node[(long long)p * L.cap + l];
    const int ip = L.dir[(long long)p * L.cap + l];
    const long long o = offset[p] + l;
    ox[o] = (int)(n % (uint32_t)g.xp) + 1;
    oy[o] = (int)((n / (uint32_t)g.xp) % (uint32_t)g.ly) + 1;
    oz[o] = (int)(n / ((uint32_t)g.xp * (uint32_t)g.ly)) + pg.globalz;
    odir[o] = ip; opart[o] = p + 1;
    int jx, jy, jz;
    link_node_coords(pg, ypglb + 3 * p, n, jx, jy, jz);
    oq[o] = link_q_rt(ypglb + 3 * p, jx, jy, jz, (R(pg.rad) * R(pg.rad)).v, ip);
}

}  // namespace d3q
