// TEST INFRASTRUCTURE ONLY -- the CUDA kernels of csrc/kernels.cuh compiled as ordinary C++ (tests/host/fake/
// cuda_runtime.h) and driven through the launch sequences of csrc/d3q19_api.cu, so that their index logic --
// the three storage phases, the wall select, the y/z wraps, 32/64-bit indices, the five-population face pack /
// unpack, the send-back after an in-place odd step, the stores into a neighbour's array of the peer-memory and
// "put" halos -- is checked bit for bit against the oracle on the GPU-less build box (tests/test_kernels_host.py).
// All z-slabs ("ranks") of a run live in this one process; what NCCL or NVLink would carry is a memcpy.
// The product never builds or loads this file; the shipped library has no CPU path.
//
// Mirrors, in d3q19_api.cu: d3q19_create (geometry), d3q19_upload_f / d3q19_download_f, exchange_faces /
// exchange_after_step, launch_step_range / launch_step_halo / launch_step_put / step_impl / step_dispatch,
// macro_launch, d3q19_vortcalc, d3q19_forcingp, d3q19_init_channel, profiles_impl, d3q19_diag, d3q19_avedensity,
// d3q19_prerelax, and the particle sequence d3q19_beads_links / _collision / _lubforce / _move / _filling /
// d3q19_particle_step (kernels with barriers, shuffles or shared memory run with one fiber per thread).
#include <cuda_runtime.h>

#include <vector>

#include "kernels.cuh"
#include "particles.cuh"

using namespace d3q;

namespace {

constexpr size_t POP_PAD = 32;
const FaceSlots SLOTS_PZ = {{5, 11, 12, 15, 16}};
const FaceSlots SLOTS_MZ = {{6, 13, 14, 17, 18}};

enum Transport { T_PACKED = 0, T_FUSED = 1, T_FUSED_SPLIT = 2, T_PUT = 3 };

struct Rank {
    Geom g;
    int globalz = 0;
    std::vector<double> A_alloc, B_alloc;
    double *A = nullptr, *B = nullptr;
    std::vector<double> rho, ux, uy, uz, ffx, ffy, ffz, vort, vhalo;
    std::vector<int32_t> solid, isn;
    std::vector<double> send_up, send_dn, recv_lo, recv_hi;
    unsigned int flags[16] = {0};
    unsigned int epoch = 0;
    // the neighbours' arrays as this rank sees them ([0] lower, [1] upper), swapped in lockstep for AB
    double *peer_A[2] = {nullptr, nullptr}, *peer_B[2] = {nullptr, nullptr};
    // particle path (per slab: masks and links; the particle tables are replicated, here per slab as on the GPUs)
    std::vector<int32_t> own;
    std::vector<double> ypglb, ypmask, wp, omgp, fHIp, torqp, flubp, forcepp, torqpp, thetap;
    std::vector<uint32_t> lnode;
    std::vector<int32_t> ldir;
    std::vector<unsigned long long> lcount;
    std::vector<uint32_t> fnode; std::vector<int32_t> fpart;      // refill list
    unsigned long long pcnt[4] = {0, 0, 0, 0};
    std::vector<double> fill_send_up, fill_send_dn, fill_ghost_lo, fill_ghost_hi;      // 19 x plane each
    unsigned long long nfilled = 0;
};

struct Sim {
    int nx, ny, nz, nranks, ab, strict, transport, idx64;
    int phase = 0;
    Mrt mrt;
    double Fx = 0, Fy = 0, Fz = 0, rho_shift = 0;
    long long pf_rows = 0;
    unsigned long long rhoerr_bits = 0;      // PRERELAX: max |rho_new - rho_old| (one word = the MAX all-reduce)
    bool want_rhoerr = false;
    bool force_field = false, has_solid = false;
    std::vector<double> ypglb, wp, omgp;      // (3,npart) tables of the solid branches (macrovar, vortcalc)
    double rhopart = 0.0;
    // particle path
    bool part_on = false, links_valid = false, mask_built = false;
    int npart = 0;
    double rad = 0, rho0 = 1, amp = 0, aip = 0, gforce[3] = {0, 0, 0};
    LubParams lub = {0, 0, 0, 0, 0, 0, 0};
    long long maxlink = 0, linkcap = 0;
    std::vector<Rank> r;
};

dim3 grid_nodes(const Geom &g, int nplanes) {
    return dim3((unsigned)((g.lx + BLOCK_X - 1) / BLOCK_X), (unsigned)g.ly, (unsigned)nplanes);
}

int read_kind(const Sim &s) { return s.ab ? READ_PULL_NAT : (s.phase == 0 ? READ_DIRECT : READ_PULL_SWAP); }

// ---- exchange_faces: pack on every rank, "send/recv", unpack on every rank ---------------------------------
// plane arguments: -1 stands for the slab's own lz, -2 for lz+1 (slabs may differ in thickness)
int plane_of(const Geom &g, int code) { return code == -1 ? g.lz : (code == -2 ? g.lz + 1 : code); }

void exchange_faces(Sim &s, bool use_B, int up_src, const FaceSlots &up_slots, int lo_dst, int dn_src,
                    const FaceSlots &dn_slots, int hi_dst, int exclude_walls) {
    const int n = s.nranks;
    for (int k = 0; k < n; ++k) {
        Rank &q = s.r[k];
        const Geom &g = q.g;
        FacePair pk;
        pk.buf[0] = q.send_up.data(); pk.zg[0] = plane_of(g, up_src); pk.slots[0] = up_slots;
        pk.buf[1] = q.send_dn.data(); pk.zg[1] = plane_of(g, dn_src); pk.slots[1] = dn_slots;
        hs_launch(dim3((unsigned)((g.xp + BLOCK_X - 1) / BLOCK_X), (unsigned)g.ly, 10u), BLOCK_X, k_face_pack, g,
                  (const double *)(use_B ? q.B : q.A), pk);
    }
    for (int k = 0; k < n; ++k) {
        const int up = (k + 1) % n, dn = (k + n - 1) % n;
        s.r[up].recv_lo = s.r[k].send_up;          // my "up" data is the upper neighbour's recv_lo
        s.r[dn].recv_hi = s.r[k].send_dn;
    }
    for (int k = 0; k < n; ++k) {
        Rank &q = s.r[k];
        const Geom &g = q.g;
        FacePair un;
        un.buf[0] = q.recv_lo.data(); un.zg[0] = plane_of(g, lo_dst); un.slots[0] = up_slots;
        un.buf[1] = q.recv_hi.data(); un.zg[1] = plane_of(g, hi_dst); un.slots[1] = dn_slots;
        hs_launch(dim3((unsigned)((g.lx + BLOCK_X - 1) / BLOCK_X), (unsigned)g.ly, 10u), BLOCK_X, k_face_unpack, g,
                  use_B ? q.B : q.A, un, exclude_walls);
    }
}
void exchange_after_step(Sim &s, int step_kind, bool use_B) {
    switch (step_kind) {
    case STEP_AB: exchange_faces(s, use_B, -1, SLOTS_PZ, 0, 1, SLOTS_MZ, -2, 0); break;
    case STEP_AA_EVEN: exchange_faces(s, use_B, -1, SLOTS_MZ, 0, 1, SLOTS_PZ, -2, 0); break;
    default: exchange_faces(s, use_B, -2, SLOTS_PZ, 1, 0, SLOTS_MZ, -1, 1); break;
    }
}

// ---- the step ------------------------------------------------------------------------------------------------
template <int SK, bool STRICT, bool GENERIC, bool HALO>
void launch_k_step(const Sim &s, const Geom &g, const StepParams &p, int nplanes) {
    if (GENERIC && p.macro_mode == 1 && p.rhoerr_bits) {       // block maximum of |rho - rhop|: shuffles + a barrier
        if (!s.idx64) hs_launch_coop(grid_nodes(g, nplanes), BLOCK_X, k_step<SK, STRICT, GENERIC, uint32_t, HALO>, p);
        else hs_launch_coop(grid_nodes(g, nplanes), BLOCK_X, k_step<SK, STRICT, GENERIC, unsigned long long, HALO>, p);
        return;
    }
    if (!s.idx64) hs_launch(grid_nodes(g, nplanes), BLOCK_X, k_step<SK, STRICT, GENERIC, uint32_t, HALO>, p);
    else hs_launch(grid_nodes(g, nplanes), BLOCK_X, k_step<SK, STRICT, GENERIC, unsigned long long, HALO>, p);
}
template <int SK, bool STRICT, bool GENERIC>
void launch_step_range(const Sim &s, const Geom &g, const StepParams &p0, int z0, int nplanes, int zstride = 1) {
    if (nplanes <= 0) return;
    StepParams p = p0;
    p.z0 = z0; p.zstride = zstride;
    launch_k_step<SK, STRICT, GENERIC, false>(s, g, p, nplanes);
}

template <int SK, bool STRICT, bool GENERIC>
void step_all_ranks(Sim &s, int macro_mode) {
    const bool ab = SK == STEP_AB;
    const int n = s.nranks;
    std::vector<StepParams> ps(n);
    for (int k = 0; k < n; ++k) {
        Rank &q = s.r[k];
        StepParams &p = ps[k];
        std::memset(&p, 0, sizeof p);
        p.g = q.g; p.mrt = s.mrt;
        p.Fx = s.Fx; p.Fy = s.Fy; p.Fz = s.Fz;
        p.rho_shift = s.rho_shift;
        const int bpr = (q.g.lx + BLOCK_X - 1) / BLOCK_X;
        p.pf_ahead = (s.pf_rows > 0 ? (s.pf_rows + bpr - 1) / bpr : 0) * q.g.xp;
        p.macro_mode = macro_mode;
        p.rho = q.rho.data(); p.ux = q.ux.data(); p.uy = q.uy.data(); p.uz = q.uz.data();
        if (s.force_field) { p.ffx = q.ffx.data(); p.ffy = q.ffy.data(); p.ffz = q.ffz.data(); }
        if (s.has_solid) p.solid = q.solid.data();
        if (s.part_on) p.solid = q.own.data() + q.g.plane;
        p.A = q.A; p.B = ab ? q.B : nullptr;
        if (s.want_rhoerr) p.rhoerr_bits = &s.rhoerr_bits;
    }
    if (n == 1) {
        launch_step_range<SK, STRICT, GENERIC>(s, s.r[0].g, ps[0], 1, s.r[0].g.lz);
    } else if (s.transport == T_PACKED) {
        for (int k = 0; k < n; ++k) {
            const int lz = s.r[k].g.lz;
            if (lz > 2) {
                launch_step_range<SK, STRICT, GENERIC>(s, s.r[k].g, ps[k], 1, 2, lz - 1);      // planes 1 and lz in one launch
                launch_step_range<SK, STRICT, GENERIC>(s, s.r[k].g, ps[k], 2, lz - 2);
            } else {
                launch_step_range<SK, STRICT, GENERIC>(s, s.r[k].g, ps[k], 1, lz);
            }
        }
        exchange_after_step(s, SK, ab);
    } else if (s.transport == T_PUT) {
        // launch_step_put: plain kernels everywhere, then every rank's face copies (flags checked below)
        for (int k = 0; k < n; ++k) {
            Rank &q = s.r[k];
            ++q.epoch;
            const int lz = q.g.lz;
            if (lz > 2) {
                launch_step_range<SK, STRICT, GENERIC>(s, q.g, ps[k], 1, 2, lz - 1);
                launch_step_range<SK, STRICT, GENERIC>(s, q.g, ps[k], 2, lz - 2);
            } else {
                launch_step_range<SK, STRICT, GENERIC>(s, q.g, ps[k], 1, lz);
            }
        }
        for (int k = 0; k < n; ++k) {
            Rank &q = s.r[k];
            const Geom &g = q.g;
            const int lz = g.lz, dn = (k + n - 1) % n, up = (k + 1) % n;
            // the copy-engine transfers of launch_step_put (d3q19_api.cu): one copy per crossing population and face, the
            // two c_x = +-1 populations of an in-place odd step as pitched copies of lx-1 columns, then the flags
            const double *src0 = ab ? q.B : q.A;
            double *dstA[2] = {ab ? q.peer_B[1] : q.peer_A[1], ab ? q.peer_B[0] : q.peer_A[0]};
            const long long slab_dst[2] = {s.r[up].g.slab, s.r[dn].g.slab};
            const int lz_dn = s.r[dn].g.lz;
            int zsrc[2], zdst[2], excl = 0;
            FaceSlots sl[2];
            switch (SK) {
            case STEP_AB: zsrc[0] = lz; sl[0] = SLOTS_PZ; zdst[0] = 0; zsrc[1] = 1; sl[1] = SLOTS_MZ; zdst[1] = lz_dn + 1; break;
            case STEP_AA_EVEN: zsrc[0] = lz; sl[0] = SLOTS_MZ; zdst[0] = 0; zsrc[1] = 1; sl[1] = SLOTS_PZ; zdst[1] = lz_dn + 1; break;
            default: zsrc[0] = lz + 1; sl[0] = SLOTS_PZ; zdst[0] = 1; zsrc[1] = 0; sl[1] = SLOTS_MZ; zdst[1] = lz_dn; excl = 1; break;
            }
            const size_t pl = (size_t)g.plane;
            for (int c = 0; c < 5; ++c)
                for (int d = 0; d < 2; ++d) {
                    const int slot = sl[d].s[c];
                    const double *src = src0 + (size_t)slot * g.slab + (size_t)zsrc[d] * pl;
                    double *dst = dstA[d] + (size_t)slot * slab_dst[d] + (size_t)zdst[d] * pl;
                    const int cx = dir_cx_rt(slot);
                    if (excl && cx != 0) {
                        const size_t off = cx > 0 ? 1 : 0;
                        for (int y = 0; y < g.ly; ++y)
                            std::memcpy(dst + off + (size_t)y * g.xp, src + off + (size_t)y * g.xp, (size_t)(g.lx - 1) * sizeof(double));
                    } else {
                        std::memcpy(dst, src, pl * sizeof(double));
                    }
                }
            hs_launch(dim3(1), 1, k_flag_raise, (unsigned int *)s.r[up].flags, (unsigned int *)(s.r[dn].flags + 1), q.epoch);
        }
    } else {
        // launch_step_halo: the stores into the neighbours happen inside the step kernel
        for (int k = 0; k < n; ++k) {
            Rank &q = s.r[k];
            StepParams p = ps[k];
            Halo &h = p.halo;
            const int dn = (k + n - 1) % n, up = (k + 1) % n;
            h.peer_dn = ab ? q.peer_B[0] : q.peer_A[0];
            h.peer_up = ab ? q.peer_B[1] : q.peer_A[1];
            h.slab_dn = s.r[dn].g.slab; h.slab_up = s.r[up].g.slab;
            h.lz_dn = s.r[dn].g.lz;
            h.wait_lo = q.flags; h.wait_hi = q.flags + 1;
            h.sig_dn = s.r[dn].flags + 1;
            h.sig_up = s.r[up].flags;
            h.ctr = q.flags + 2;
            h.err = q.flags + 8;
            h.timeout_ns = 1000;
            h.epoch = ++q.epoch;
            const bool split = s.transport == T_FUSED_SPLIT && q.g.lz > 2;
            const dim3 gr = grid_nodes(q.g, split ? 2 : q.g.lz);
            h.nblk_face = gr.x * gr.y;
            launch_k_step<SK, STRICT, GENERIC, true>(s, q.g, p, split ? 2 : q.g.lz);
            if (split) launch_step_range<SK, STRICT, GENERIC>(s, q.g, ps[k], 2, q.g.lz - 2);
        }
    }
    if (n > 1 && s.transport != T_PACKED) {
        // every flag a neighbour had to raise in this step is up, nobody hit the watchdog
        for (int k = 0; k < n; ++k) {
            Rank &q = s.r[k];
            if (q.flags[0] != q.epoch || q.flags[1] != q.epoch || q.flags[8] != 0u || q.flags[2] != 0u || q.flags[3] != 0u) {
                std::fprintf(stderr, "host kernel harness: halo flags of slab %d after step %u: %u %u ctr %u %u err %u\n", k,
                             q.epoch, q.flags[0], q.flags[1], q.flags[2], q.flags[3], q.flags[8]);
                std::abort();
            }
        }
    }
    if (ab) {
        for (Rank &q : s.r) {
            std::swap(q.A, q.B);
            std::swap(q.peer_A[0], q.peer_B[0]);
            std::swap(q.peer_A[1], q.peer_B[1]);
        }
    }
}

template <bool STRICT, bool GENERIC>
void step_dispatch(Sim &s, int macro_mode) {
    if (s.ab) {
        step_all_ranks<STEP_AB, STRICT, GENERIC>(s, macro_mode);
    } else if (s.phase == 0) {
        step_all_ranks<STEP_AA_EVEN, STRICT, GENERIC>(s, macro_mode);
        s.phase = 1;
    } else {
        step_all_ranks<STEP_AA_ODD, STRICT, GENERIC>(s, macro_mode);
        s.phase = 0;
    }
}

void collide_stream(Sim &s, int macro_mode) {
    const bool generic = macro_mode != 0 || s.force_field || s.has_solid || s.part_on || s.rho_shift != 0.0;
    if (generic) { if (s.strict) step_dispatch<true, true>(s, macro_mode); else step_dispatch<false, true>(s, macro_mode); }
    else { if (s.strict) step_dispatch<true, false>(s, macro_mode); else step_dispatch<false, false>(s, macro_mode); }
    if (macro_mode == 0) s.rho_shift = 0.0;
}

template <int RK>
void macro_rank(const Sim &s, Rank &q, int rho_only) {
    MacroParams p;
    std::memset(&p, 0, sizeof p);
    p.g = q.g; p.A = q.A;
    p.rho = q.rho.data(); p.ux = q.ux.data(); p.uy = q.uy.data(); p.uz = q.uz.data();
    p.Fx = s.Fx; p.Fy = s.Fy; p.Fz = s.Fz;
    if (s.force_field) { p.ffx = q.ffx.data(); p.ffy = q.ffy.data(); p.ffz = q.ffz.data(); }
    if (s.has_solid) { p.solid = q.solid.data(); p.isnodes = q.isn.data(); }
    if (!s.ypglb.empty()) { p.ipart = 1; p.ypglb = s.ypglb.data(); p.wp = s.wp.data(); p.omgp = s.omgp.data(); p.rhopart = s.rhopart; }
    p.ny = s.ny; p.nz = s.nz; p.globalz = q.globalz;
    p.rho_only = rho_only;
    hs_launch(grid_nodes(q.g, q.g.lz), BLOCK_X, k_macro<RK>, p);
}

template <int RK>
void gather_rank(Rank &q, double *aos) {
    hs_launch(grid_nodes(q.g, q.g.lz), BLOCK_X, k_gather_aos<RK>, q.g, (const double *)q.A, aos, 1);
}

template <int RK>
void profiles_rank(const Sim &s, Rank &q, int rows_per_chunk, int nchunks, double *partial, double *out) {
    hs_launch(dim3((unsigned)((q.g.lx + BLOCK_X - 1) / BLOCK_X), (unsigned)nchunks, 1u), BLOCK_X, k_profiles<RK>, q.g,
              (const double *)q.A, s.Fx, s.Fy, s.Fz, s.force_field ? (const double *)q.ffx.data() : (const double *)nullptr,
              s.force_field ? (const double *)q.ffy.data() : (const double *)nullptr,
              s.force_field ? (const double *)q.ffz.data() : (const double *)nullptr,
              s.has_solid ? (const int32_t *)q.solid.data() : (const int32_t *)nullptr, rows_per_chunk, partial);
    hs_launch(dim3((unsigned)((q.g.lx + BLOCK_X - 1) / BLOCK_X), (unsigned)NPROF, 1u), BLOCK_X, k_profiles_final, q.g.lx, nchunks,
              (const double *)partial, out);
}


PartGeom part_geom(const Sim &s, const Rank &q) {
    PartGeom pg;
    pg.g = q.g; pg.nx = s.nx; pg.ny = s.ny; pg.nz = s.nz; pg.globalz = q.globalz; pg.rad = s.rad;
    return pg;
}

Links links_of(const Sim &s, Rank &q) { return Links{q.lnode.data(), q.ldir.data(), q.lcount.data(), s.linkcap}; }

static inline dim3 sweep_grid(const Sim &s) {
    return dim3((unsigned)s.npart, (unsigned)((part_max_rows(s.rad) + PART_WARPS - 1) / PART_WARPS));
}
static FillList fill_of(Rank &q, long long cap) { return FillList{q.fnode.data(), q.fpart.data(), &q.pcnt[1], cap}; }

// d3q19_beads_links on every slab
void beads_links(Sim &s) {
    for (Rank &q : s.r) {
        const PartGeom pg = part_geom(s, q);
        const dim3 gs = sweep_grid(s);
        q.pcnt[1] = 0;
        std::fill(q.lcount.begin(), q.lcount.end(), 0ull);
        if (s.mask_built)
            hs_launch(gs, 32 * PART_WARPS, k_beads_uncover, pg, s.npart, (const double *)q.ypmask.data(), (const double *)q.ypglb.data(),
                      q.own.data(), fill_of(q, s.maxlink));
        hs_launch(gs, 32 * PART_WARPS, k_beads_cover, pg, s.npart, (const double *)q.ypglb.data(), q.own.data());
        q.ypmask = q.ypglb;
        hs_launch_coop(gs, 32 * PART_WARPS, k_beads_links, pg, s.npart, (const double *)q.ypglb.data(), (const int32_t *)q.own.data(),
                       links_of(s, q));
        for (int p = 0; p < s.npart; ++p)
            if ((long long)q.lcount[p] > s.linkcap) hs::trap("more links than a particle's segment holds");
    }
    s.mask_built = true;
    s.links_valid = true;
}

template <int RK>
void ibb_rank(Sim &s, Rank &q) {
    IbbParams P;
    P.pg = part_geom(s, q); P.S = q.A; P.own = q.own.data(); P.L = links_of(s, q);
    P.ypglb = q.ypglb.data(); P.wp = q.wp.data(); P.omgp = q.omgp.data(); P.rho0 = s.rho0;
    P.fHIp = q.fHIp.data(); P.torqp = q.torqp.data();
    // like the device path: one grid row per particle, wide enough for the longest segment in use plus idle blocks
    unsigned long long longest = 0;
    for (int p = 0; p < s.npart; ++p) longest = q.lcount[p] > longest ? q.lcount[p] : longest;
    hs_launch_coop(dim3((unsigned)((longest + 300 + 127) / 128), (unsigned)s.npart), 128, k_beads_ibb<RK>, P);
}
void beads_collision(Sim &s) {
    for (Rank &q : s.r) {
        std::fill(q.fHIp.begin(), q.fHIp.end(), 0.0);
        std::fill(q.torqp.begin(), q.torqp.end(), 0.0);
        switch (read_kind(s)) {
        case READ_DIRECT: ibb_rank<READ_DIRECT>(s, q); break;
        case READ_PULL_NAT: ibb_rank<READ_PULL_NAT>(s, q); break;
        default: ibb_rank<READ_PULL_SWAP>(s, q); break;
        }
    }
    if (s.nranks > 1) {          // ncclAllReduce(SUM) of (fHIp, torqp) over the slabs, in rank order
        std::vector<double> f(3 * s.npart, 0.0), t(3 * s.npart, 0.0);
        for (Rank &q : s.r)
            for (int i = 0; i < 3 * s.npart; ++i) { f[i] += q.fHIp[i]; t[i] += q.torqp[i]; }
        for (Rank &q : s.r) { q.fHIp = f; q.torqp = t; }
    }
}
static void lubmove(Sim &s, int do_lub, int do_move) {
    for (Rank &q : s.r) {
        MoveParams M = {s.amp, s.aip, s.gforce[0], s.gforce[1], s.gforce[2], q.fHIp.data(), q.torqp.data(), q.flubp.data(),
                        q.forcepp.data(), q.torqpp.data(), q.ypglb.data(), q.wp.data(), q.omgp.data(), q.thetap.data()};
        const int nt = s.npart < 32 ? s.npart * 32 : 1024;
        hs_launch_coop(dim3(1), nt, k_beads_lubmove, part_geom(s, q), s.npart, (const double *)q.ypglb.data(), s.lub, q.flubp.data(), M,
                       do_lub, do_move);
    }
    if (do_move) s.links_valid = false;
}
void beads_lubforce(Sim &s) { lubmove(s, 1, 0); }
void beads_move(Sim &s) { lubmove(s, 0, 1); }
template <int RK>
void fill_rank(Sim &s, Rank &q) {
    FillParams P;
    P.pg = part_geom(s, q); P.S = q.A; P.own = q.own.data(); P.F = fill_of(q, s.maxlink);
    P.ypglb = q.ypglb.data(); P.wp = q.wp.data(); P.omgp = q.omgp.data(); P.nfilled = &q.nfilled;
    P.ghost_lo = P.ghost_hi = nullptr;
    if (s.nranks > 1) { P.ghost_lo = q.fill_ghost_lo.data(); P.ghost_hi = q.fill_ghost_hi.data(); }
    hs_launch(dim3(3), 128, k_beads_fill<RK>, P);                      // a fixed grid strides over the list
    q.pcnt[1] = 0;
}
template <int RK>
void plane_gather_rank(Rank &q) {
    const size_t cnt = (size_t)NPOP * q.g.plane;
    q.fill_send_up.assign(cnt, 0.0); q.fill_send_dn.assign(cnt, 0.0);
    hs_launch(grid_nodes(q.g, 2), BLOCK_X, k_plane_gather<RK>, q.g, (const double *)q.A, q.fill_send_up.data(), q.g.lz,
              q.fill_send_dn.data(), 1);
}
void beads_filling(Sim &s) {
    if (s.nranks > 1) {
        for (Rank &q : s.r) {
            switch (read_kind(s)) {
            case READ_DIRECT: plane_gather_rank<READ_DIRECT>(q); break;
            case READ_PULL_NAT: plane_gather_rank<READ_PULL_NAT>(q); break;
            default: plane_gather_rank<READ_PULL_SWAP>(q); break;
            }
        }
        const int n = s.nranks;
        for (int k = 0; k < n; ++k) {
            s.r[(k + 1) % n].fill_ghost_lo = s.r[k].fill_send_up;
            s.r[(k + n - 1) % n].fill_ghost_hi = s.r[k].fill_send_dn;
        }
    }
    for (Rank &q : s.r) {
        q.nfilled = 0;
        switch (read_kind(s)) {
        case READ_DIRECT: fill_rank<READ_DIRECT>(s, q); break;
        case READ_PULL_NAT: fill_rank<READ_PULL_NAT>(s, q); break;
        default: fill_rank<READ_PULL_SWAP>(s, q); break;
        }
    }
}

template <int RK>
void diag_rank(const Sim &s, Rank &q, int rows_per_chunk, std::vector<double> &out) {
    const long long nrows = (long long)q.g.ly * q.g.lz;
    const int nchunks = (int)((nrows + rows_per_chunk - 1) / rows_per_chunk);
    const dim3 gd((unsigned)((q.g.lx + BLOCK_X - 1) / BLOCK_X), (unsigned)nchunks, 1u);
    std::vector<double> partial((size_t)gd.x * gd.y * NDIAG, 0.0);
    const int32_t *solid = s.part_on ? q.own.data() + q.g.plane : (s.has_solid ? q.solid.data() : nullptr);
    const double *ff[3] = {s.force_field ? q.ffx.data() : nullptr, s.force_field ? q.ffy.data() : nullptr,
                           s.force_field ? q.ffz.data() : nullptr};
    hs_launch_coop(gd, BLOCK_X, k_diag<RK>, q.g, (const double *)q.A, s.Fx, s.Fy, s.Fz, ff[0], ff[1], ff[2], solid, rows_per_chunk,
                   partial.data());
    out.assign(NDIAG, 0.0);
    hs_launch(dim3(1), 1, k_diag_final, (int)(gd.x * gd.y), (const double *)partial.data(), out.data());
}

}  // namespace

extern "C" {

// lz[k]: thickness of slab k (para.f90:241-245 decides; the caller passes it).  scheme: 0 AA, 1 AB.
void *hs_create(int nx, int ny, int nz, int nranks, const int *lz, int scheme_ab, int strict, int transport, int idx64,
                const double *mrt10, double Fx, double Fy, double Fz, int pf_blocks) {
    Sim *s = new Sim();
    s->nx = nx; s->ny = ny; s->nz = nz; s->nranks = nranks; s->ab = scheme_ab; s->strict = strict;
    s->transport = transport; s->idx64 = idx64;
    s->mrt = Mrt{mrt10[0], mrt10[1], mrt10[2], mrt10[3], mrt10[4], mrt10[5], mrt10[6], mrt10[7], mrt10[8], mrt10[9]};
    s->Fx = Fx; s->Fy = Fy; s->Fz = Fz;
    s->pf_rows = pf_blocks;
    s->r.resize(nranks);
    int gz = 0;
    for (int k = 0; k < nranks; ++k) {
        Rank &q = s->r[k];
        Geom &g = q.g;
        g.lx = nx; g.ly = ny; g.lz = lz[k];
        g.xp = (nx + 15) / 16 * 16;
        g.plane = (long long)g.xp * g.ly;
        g.slab = g.plane * (g.lz + 2);
        g.zlo_src = nranks == 1 ? g.lz : 0;
        g.zhi_src = nranks == 1 ? 1 : g.lz + 1;
        q.globalz = gz; gz += lz[k];
        const size_t n = (size_t)NPOP * g.slab + 2 * POP_PAD;
        // poison instead of zeros: a read of a never-written element must not pass unnoticed
        q.A_alloc.assign(n, std::nan(""));
        q.A = q.A_alloc.data() + POP_PAD;
        if (scheme_ab) { q.B_alloc.assign(n, std::nan("")); q.B = q.B_alloc.data() + POP_PAD; }
        const size_t nf = (size_t)g.plane * g.lz;
        q.rho.assign(nf, 0.0); q.ux.assign(nf, 0.0); q.uy.assign(nf, 0.0); q.uz.assign(nf, 0.0);
        q.send_up.assign(5 * g.plane, 0.0); q.send_dn.assign(5 * g.plane, 0.0);
        q.recv_lo.assign(5 * g.plane, 0.0); q.recv_hi.assign(5 * g.plane, 0.0);
    }
    for (int k = 0; k < nranks; ++k) {
        const int dn = (k + nranks - 1) % nranks, up = (k + 1) % nranks;
        s->r[k].peer_A[0] = s->r[dn].A; s->r[k].peer_B[0] = s->r[dn].B;
        s->r[k].peer_A[1] = s->r[up].A; s->r[k].peer_B[1] = s->r[up].B;
    }
    return s;
}

void hs_destroy(void *h) { delete (Sim *)h; }

// d3q19_upload_f for every slab; f is the global canonical AoS f[iz][iy][ix][ip]
void hs_upload(void *h, const double *f) {
    Sim &s = *(Sim *)h;
    const size_t per_plane = (size_t)NPOP * s.nx * s.ny;
    for (Rank &q : s.r)
        hs_launch(grid_nodes(q.g, q.g.lz), BLOCK_X, k_scatter_aos, q.g, s.ab ? q.B : q.A, f + (size_t)q.globalz * per_plane, 1);
    if (s.ab) {
        if (s.nranks > 1) exchange_faces(s, true, -1, SLOTS_MZ, 0, 1, SLOTS_PZ, -2, 0);
        for (Rank &q : s.r) hs_launch(grid_nodes(q.g, q.g.lz), BLOCK_X, k_unstream, q.g, (const double *)q.B, q.A);
        if (s.nranks > 1) exchange_after_step(s, STEP_AB, false);
    }
    s.phase = 0;
}

void hs_download(void *h, double *f) {
    Sim &s = *(Sim *)h;
    const size_t per_plane = (size_t)NPOP * s.nx * s.ny;
    for (Rank &q : s.r) {
        double *o = f + (size_t)q.globalz * per_plane;
        switch (read_kind(s)) {
        case READ_DIRECT: gather_rank<READ_DIRECT>(q, o); break;
        case READ_PULL_NAT: gather_rank<READ_PULL_NAT>(q, o); break;
        default: gather_rank<READ_PULL_SWAP>(q, o); break;
        }
    }
}

void hs_steps(void *h, int n, int macro_mode) {
    Sim &s = *(Sim *)h;
    for (int i = 0; i < n; ++i) collide_stream(s, macro_mode);
}

void hs_set_rho_shift(void *h, double v) { ((Sim *)h)->rho_shift = v; }

// pitched device field <-> global host layout a[iz][iy][ix]
static void field_in(const Sim &s, const Rank &q, std::vector<double> &dev, const double *host) {
    dev.assign((size_t)q.g.plane * q.g.lz, 0.0);
    hs_launch(dim3((unsigned)((q.g.lx + BLOCK_X - 1) / BLOCK_X), (unsigned)q.g.ly, (unsigned)q.g.lz), BLOCK_X, k_field_unpack, q.g.lx,
              q.g.xp, dev.data(), host + (size_t)q.globalz * s.nx * s.ny);
}
static void field_out(const Sim &s, const Rank &q, const double *dev, double *host) {
    hs_launch(dim3((unsigned)((q.g.lx + BLOCK_X - 1) / BLOCK_X), (unsigned)q.g.ly, (unsigned)q.g.lz), BLOCK_X, k_field_pack, q.g.lx,
              q.g.xp, dev, host + (size_t)q.globalz * s.nx * s.ny, 1);
}

void hs_set_macro(void *h, const double *rho, const double *ux, const double *uy, const double *uz) {
    Sim &s = *(Sim *)h;
    for (Rank &q : s.r) { field_in(s, q, q.rho, rho); field_in(s, q, q.ux, ux); field_in(s, q, q.uy, uy); field_in(s, q, q.uz, uz); }
}
void hs_set_force_field(void *h, const double *fx, const double *fy, const double *fz) {
    Sim &s = *(Sim *)h;
    for (Rank &q : s.r) { field_in(s, q, q.ffx, fx); field_in(s, q, q.ffy, fy); field_in(s, q, q.ffz, fz); }
    s.force_field = true;
}
// ibnodes as the global un-ghosted (nz,ny,nx) array: -1 fluid, > 0 solid; isnodes the owner (1-based)
void hs_set_solid(void *h, const int32_t *ib, const int32_t *isn, int npart, const double *ypglb, const double *wp,
                  const double *omgp, double rhopart) {
    Sim &s = *(Sim *)h;
    if (npart > 0) {
        s.ypglb.assign(ypglb, ypglb + 3 * npart); s.wp.assign(wp, wp + 3 * npart); s.omgp.assign(omgp, omgp + 3 * npart);
        s.rhopart = rhopart;
    }
    for (Rank &q : s.r) {
        const Geom &g = q.g;
        q.solid.assign((size_t)g.plane * g.lz, -1);
        q.isn.assign((size_t)g.plane * g.lz, 0);
        const dim3 gr((unsigned)((g.lx + BLOCK_X - 1) / BLOCK_X), (unsigned)g.ly, (unsigned)g.lz);
        hs_launch(gr, BLOCK_X, k_field_unpack_i32, g.lx, g.xp, q.solid.data(), ib + (size_t)q.globalz * s.nx * s.ny, 0, g.ly);
        if (isn) hs_launch(gr, BLOCK_X, k_field_unpack_i32, g.lx, g.xp, q.isn.data(), isn + (size_t)q.globalz * s.nx * s.ny, 0, g.ly);
    }
    s.has_solid = true;
}

// macrovar (rho_only = 0) / rhoupdat (rho_only = 1), then the four fields in the global host layout
void hs_macrovar(void *h, int rho_only, double *rho, double *ux, double *uy, double *uz) {
    Sim &s = *(Sim *)h;
    for (Rank &q : s.r) {
        switch (read_kind(s)) {
        case READ_DIRECT: macro_rank<READ_DIRECT>(s, q, rho_only); break;
        case READ_PULL_NAT: macro_rank<READ_PULL_NAT>(s, q, rho_only); break;
        default: macro_rank<READ_PULL_SWAP>(s, q, rho_only); break;
        }
        field_out(s, q, q.rho.data(), rho); field_out(s, q, q.ux.data(), ux);
        field_out(s, q, q.uy.data(), uy); field_out(s, q, q.uz.data(), uz);
    }
}

// d3q19_vortcalc: velocity planes next to the faces travel like exchng8's z phase
void hs_vortcalc(void *h, double *ox, double *oy, double *oz) {
    Sim &s = *(Sim *)h;
    const int n = s.nranks;
    for (Rank &q : s.r) { q.vort.assign(3 * (size_t)q.g.plane * q.g.lz, 0.0); q.vhalo.assign(6 * (size_t)q.g.plane, 0.0); }
    for (int k = 0; k < n && n > 1; ++k) {
        Rank &q = s.r[k];
        const size_t pl = (size_t)q.g.plane;
        const Rank &dn = s.r[(k + n - 1) % n], &up = s.r[(k + 1) % n];
        const double *src_dn[3] = {dn.ux.data(), dn.uy.data(), dn.uz.data()};
        const double *src_up[3] = {up.ux.data(), up.uy.data(), up.uz.data()};
        for (int c = 0; c < 3; ++c) {
            std::memcpy(q.vhalo.data() + c * pl, src_dn[c] + (size_t)(dn.g.lz - 1) * pl, pl * sizeof(double));     // zlo
            std::memcpy(q.vhalo.data() + (3 + c) * pl, src_up[c], pl * sizeof(double));                             // zhi
        }
    }
    for (Rank &q : s.r) {
        const size_t nf = (size_t)q.g.plane * q.g.lz;
        VortParams p;
        std::memset(&p, 0, sizeof p);
        p.g = q.g; p.ux = q.ux.data(); p.uy = q.uy.data(); p.uz = q.uz.data();
        p.ox = q.vort.data(); p.oy = q.vort.data() + nf; p.oz = q.vort.data() + 2 * nf;
        if (n > 1) { p.zlo = q.vhalo.data(); p.zhi = q.vhalo.data() + 3 * q.g.plane; }
        if (s.has_solid && !s.omgp.empty()) { p.solid = q.solid.data(); p.isnodes = q.isn.data(); p.omgp = s.omgp.data(); }
        hs_launch(grid_nodes(q.g, q.g.lz), BLOCK_X, k_vortcalc, p);
        field_out(s, q, p.ox, ox); field_out(s, q, p.oy, oy); field_out(s, q, p.oz, oz);
    }
}

// d3q19_profiles: out[12][nx] summed over the slabs in rank order
void hs_profiles(void *h, int rows_per_chunk, double *out) {
    Sim &s = *(Sim *)h;
    std::vector<double> acc((size_t)NPROF * s.nx, 0.0);
    for (Rank &q : s.r) {
        const long long nrows = (long long)q.g.ly * q.g.lz;
        const int nchunks = (int)((nrows + rows_per_chunk - 1) / rows_per_chunk);
        std::vector<double> partial((size_t)nchunks * NPROF * q.g.lx, 0.0), o((size_t)NPROF * q.g.lx, 0.0);
        switch (read_kind(s)) {
        case READ_DIRECT: profiles_rank<READ_DIRECT>(s, q, rows_per_chunk, nchunks, partial.data(), o.data()); break;
        case READ_PULL_NAT: profiles_rank<READ_PULL_NAT>(s, q, rows_per_chunk, nchunks, partial.data(), o.data()); break;
        default: profiles_rank<READ_PULL_SWAP>(s, q, rows_per_chunk, nchunks, partial.data(), o.data()); break;
        }
        for (size_t i = 0; i < acc.size(); ++i) acc[i] += o[i];
    }
    std::memcpy(out, acc.data(), acc.size() * sizeof(double));
}

// d3q19_init_channel on every slab (AB: un-streamed storage + the ghost fill of a step)
void hs_init_channel(void *h, double ustar, double ystar, double A9, double noise_amp, unsigned long long seed, int ivel) {
    Sim &s = *(Sim *)h;
    for (Rank &q : s.r) {
        InitParams p;
        std::memset(&p, 0, sizeof p);
        p.g = q.g; p.A = q.A;
        p.nx = s.nx; p.ny = s.ny; p.nz = s.nz; p.globalz = q.globalz;
        p.ustar = ustar; p.ystar = ystar; p.A9 = A9; p.noise_amp = noise_amp;
        p.pi2 = 2.0 * (4.0 * std::atan(1.0));
        p.seed = seed; p.ivel = ivel; p.unstream = s.ab;
        hs_launch(grid_nodes(q.g, q.g.lz), BLOCK_X, k_init_channel, p);
    }
    s.phase = 0;
    if (s.ab && s.nranks > 1) exchange_after_step(s, STEP_AB, false);
}

// d3q19_forcingp: the perturbation force field of step istep on every slab -> global host layout
void hs_forcingp(void *h, int ihh, int ixs0, double force_in_y, double Amp0, double beta9, double gamma9, double phase9,
                 double *fx, double *fy, double *fz) {
    Sim &s = *(Sim *)h;
    for (Rank &q : s.r) {
        const size_t nf = (size_t)q.g.plane * q.g.lz;
        q.ffx.assign(nf, 0.0); q.ffy.assign(nf, 0.0); q.ffz.assign(nf, 0.0);
        ForcingpParams p;
        std::memset(&p, 0, sizeof p);
        p.lx = q.g.lx; p.ly = q.g.ly; p.lz = q.g.lz; p.xp = q.g.xp;
        p.nx = s.nx; p.ny = s.ny; p.nz = s.nz; p.globalz = q.globalz;
        p.ihh = ihh; p.ixs0 = ixs0;
        p.force_in_y = force_in_y; p.Amp0 = Amp0; p.beta9 = beta9; p.gamma9 = gamma9; p.phase9 = phase9;
        p.pi2 = 2.0 * (4.0 * std::atan(1.0));
        p.fx = q.ffx.data(); p.fy = q.ffy.data(); p.fz = q.ffz.data();
        hs_launch(grid_nodes(q.g, q.g.lz), BLOCK_X, k_forcingp, p);
        field_out(s, q, p.fx, fx); field_out(s, q, p.fy, fy); field_out(s, q, p.fz, fz);
    }
    s.force_field = true;
}

// ---- reductions ---------------------------------------------------------------------------------------------------
// d3q19_avedensity: two-stage fixed-order sum per slab, all-reduce in rank order, shift of every node (collision.f90:497-511)
double hs_avedensity(void *h, long long *nfluid) {
    Sim &s = *(Sim *)h;
    double sum = 0.0;
    long long cnt = 0;
    for (Rank &q : s.r) {
        const long long nrows_ = (long long)q.g.ly * q.g.lz;
        const int nblk = nrows_ < 1024 ? (int)nrows_ : 1024;          // d3q19_avedensity: a block per (y,z) row at most
        std::vector<double> pd(nblk, 0.0);
        std::vector<long long> pc(nblk + 8, 0);
        double s1 = 0.0;
        long long c1 = 0;
        const int32_t *solid = s.part_on ? q.own.data() + q.g.plane : (s.has_solid ? q.solid.data() : nullptr);
        hs_launch_coop(dim3(nblk), 256, k_rho_partial, q.g.lx, q.g.xp, (long long)q.g.ly * q.g.lz, (const double *)q.rho.data(),
                       solid, pd.data(), pc.data());
        hs_launch(dim3(1), 32, k_rho_final, nblk, (const double *)pd.data(), (const long long *)pc.data(), &s1, &c1);
        sum += s1; cnt += c1;
    }
    const double mean = sum / (double)cnt;
    for (Rank &q : s.r)
        hs_launch(grid_nodes(q.g, q.g.lz), BLOCK_X, k_rho_shift, q.g.lx, q.g.xp, q.rho.data(), (const double *)&mean);
    s.rho_shift += mean;
    if (nfluid) *nfluid = cnt;
    return mean;
}

// d3q19_prerelax (main.f90:70-90): fused rhoupdat + collision with frozen u, max |rho - rhop| over all slabs
int hs_prerelax(void *h, double tol, int maxiter, double *rhoerrmax) {
    Sim &s = *(Sim *)h;
    int istep = 0;
    double err = 0.0;
    for (;;) {
        unsigned long long bits = 0;
        s.rhoerr_bits = 0;                 // one word shared by all slabs = the all-reduce(MAX) of the per-slab words
        s.want_rhoerr = true;
        if (s.strict) {
            if (s.ab) step_all_ranks<STEP_AB, true, true>(s, 1);
            else if (s.phase == 0) { step_all_ranks<STEP_AA_EVEN, true, true>(s, 1); s.phase = 1; }
            else { step_all_ranks<STEP_AA_ODD, true, true>(s, 1); s.phase = 0; }
        } else {
            if (s.ab) step_all_ranks<STEP_AB, false, true>(s, 1);
            else if (s.phase == 0) { step_all_ranks<STEP_AA_EVEN, false, true>(s, 1); s.phase = 1; }
            else { step_all_ranks<STEP_AA_ODD, false, true>(s, 1); s.phase = 0; }
        }
        s.want_rhoerr = false;
        bits = s.rhoerr_bits;
        std::memcpy(&err, &bits, sizeof err);
        if (err <= tol || istep > maxiter) break;
        istep++;
    }
    if (rhoerrmax) *rhoerrmax = err;
    return istep;
}

// d3q19_diag: out[12] per slab merged like the host side of d3q19_diag does (first occurrence of the maximum wins)
void hs_diag(void *h, int rows_per_chunk, double *out12_per_rank) {
    Sim &s = *(Sim *)h;
    for (int k = 0; k < s.nranks; ++k) {
        std::vector<double> o;
        switch (read_kind(s)) {
        case READ_DIRECT: diag_rank<READ_DIRECT>(s, s.r[k], rows_per_chunk, o); break;
        case READ_PULL_NAT: diag_rank<READ_PULL_NAT>(s, s.r[k], rows_per_chunk, o); break;
        default: diag_rank<READ_PULL_SWAP>(s, s.r[k], rows_per_chunk, o); break;
        }
        std::memcpy(out12_per_rank + (size_t)k * NDIAG, o.data(), NDIAG * sizeof(double));
    }
}

// ---- particle path --------------------------------------------------------------------------------------------------
void hs_particles_init(void *h, int npart, double rad, double rho0, double rhopart, const double *lub7, const double *gforce,
                       const double *ypglb, const double *wp, const double *omgp) {
    Sim &s = *(Sim *)h;
    s.part_on = true; s.npart = npart; s.rad = rad; s.rho0 = rho0;
    s.lub = LubParams{lub7[0], lub7[1], lub7[2], lub7[3], lub7[4], lub7[5], lub7[6]};
    for (int d = 0; d < 3; ++d) s.gforce[d] = gforce[d];
    const double pi = 4.0 * std::atan(1.0);
    s.linkcap = (long long)(8.0 * 4.0 * pi * (rad + 1.0) * (rad + 1.0)) + 64;
    s.maxlink = s.linkcap * npart;
    const double volp = 4.0 / 3.0 * pi * rad * rad * rad;
    s.amp = rhopart * volp;
    s.aip = 0.4 * s.amp * rad * rad;
    s.ypglb.assign(ypglb, ypglb + 3 * npart); s.wp.assign(wp, wp + 3 * npart); s.omgp.assign(omgp, omgp + 3 * npart);
    s.rhopart = rhopart;
    for (Rank &q : s.r) {
        const size_t nown = (size_t)q.g.plane * (q.g.lz + 2), tb = (size_t)3 * npart;
        q.own.assign(nown, -1);
        q.ypglb.assign(ypglb, ypglb + tb); q.ypmask = q.ypglb;
        q.wp.assign(wp, wp + tb); q.omgp.assign(omgp, omgp + tb);
        for (std::vector<double> *v : {&q.fHIp, &q.torqp, &q.flubp, &q.forcepp, &q.torqpp, &q.thetap}) v->assign(tb, 0.0);
        q.lnode.assign(s.maxlink, 0u); q.ldir.assign(s.maxlink, 0); q.lcount.assign(npart, 0ull);
        q.fnode.assign(s.maxlink, 0u); q.fpart.assign(s.maxlink, 0);
    }
}

long long hs_beads_links(void *h) {
    Sim &s = *(Sim *)h;
    beads_links(s);
    long long n = 0;
    for (Rank &q : s.r)
        for (int p = 0; p < s.npart; ++p) n += (long long)q.lcount[p];
    return n;
}

void hs_particle_step(void *h, int move) {
    Sim &s = *(Sim *)h;
    if (!s.links_valid) beads_links(s);
    collide_stream(s, 0);
    beads_collision(s);
    if (move) {
        beads_lubforce(s);
        beads_move(s);
        beads_links(s);
        beads_filling(s);
    }
}

// global owner mask own[iz][iy][ix]
void hs_get_mask(void *h, int32_t *own) {
    Sim &s = *(Sim *)h;
    for (Rank &q : s.r)
        for (int z = 0; z < q.g.lz; ++z)
            for (int y = 0; y < q.g.ly; ++y)
                for (int x = 0; x < s.nx; ++x) {          // "uncovered in the last update" (-(q+2)) is fluid to the caller
                    const int32_t v = q.own[(size_t)q.g.plane * (z + 1) + (size_t)q.g.xp * y + x];
                    own[((size_t)(q.globalz + z) * s.ny + y) * s.nx + x] = v < 0 ? -1 : v;
                }
}

// links of slab k in its order (global 1-based node coordinates); returns the count
long long hs_get_links(void *h, int k, int32_t *x, int32_t *y, int32_t *z, int32_t *ip, int32_t *part, double *qv) {
    Sim &s = *(Sim *)h;
    Rank &q = s.r[k];
    std::vector<long long> off((size_t)s.npart + 1, 0);
    for (int p = 0; p < s.npart; ++p) off[p + 1] = off[p] + (long long)q.lcount[p];
    const long long n = off[s.npart];
    if (n > 0 && x)
        hs_launch(dim3((unsigned)((s.linkcap + 127) / 128), (unsigned)s.npart), 128, k_links_export, part_geom(s, q), links_of(s, q),
                  (const double *)q.ypglb.data(), (const long long *)off.data(), x, y, z, ip, part, qv);
    return n;
}

void hs_get_particles(void *h, double *ypglb, double *wp, double *omgp, double *fHIp, double *torqp) {
    Sim &s = *(Sim *)h;
    const Rank &q = s.r[0];
    const size_t tb = (size_t)3 * s.npart * sizeof(double);
    std::memcpy(ypglb, q.ypglb.data(), tb); std::memcpy(wp, q.wp.data(), tb); std::memcpy(omgp, q.omgp.data(), tb);
    std::memcpy(fHIp, q.fHIp.data(), tb); std::memcpy(torqp, q.torqp.data(), tb);
}

}  // extern "C"
