/*
 * d3q19_b200.h -- C-ABI of libd3q19b200.so: the B200 (sm_100a) implementation of the
 * UDel-CFD D3Q19 Channel-Flow time-step hot path.
 *
 * This is the drop-in boundary.  The reference's time loop (main.f90:142-208) calls the
 * argument-less Fortran subroutines of collision.f90, which talk through `module var_inc`.
 * A replacement collision.f90 (d3q19-single-phase_b200/fortran/collision_b200.f90) keeps
 * those subroutine names and forwards each to one entry point below through ISO_C_BINDING;
 * INTEGRATION.md shows the binding.  Reference interface replaced (file:line relative to
 * /root/reference/Channel-Flow/):
 *
 *   collision_MRT  collision.f90:24-273 (+ collisionExchnge :281-372) -> d3q19_collide_stream / d3q19_shim_collision_mrt
 *   macrovar       collision.f90:378-463                              -> d3q19_macrovar + d3q19_download_macro / d3q19_shim_macrovar
 *   rhoupdat       collision.f90:469-480                              -> d3q19_rhoupdat / d3q19_shim_rhoupdat
 *   avedensity     collision.f90:487-513                              -> d3q19_avedensity / d3q19_shim_avedensity
 *   FORCING        collision.f90:515-527                              -> d3q19_set_force_uniform / d3q19_shim_forcing
 *   FORCINGP       collision.f90:529-602 (force arrays)               -> d3q19_forcingp (device) / d3q19_set_force_field (host arrays)
 *   allocarray     para.f90:418-503 (f, rho, u, ibnodes)              -> d3q19_create (device-side twins)
 *   MPI_ISEND/IRECV collision.f90:309-314,351-356                     -> NCCL send/recv inside d3q19_collide_stream
 *   MPI_ALLREDUCE  collision.f90:500-501                              -> ncclAllReduce inside d3q19_avedensity
 *
 * Conventions: plain C types only (double = Fortran real under -r8, int32_t = integer);
 * every function returns 0 on success, nonzero on failure with a message available from
 * d3q19_last_error(); no exceptions cross the ABI.  One host thread per handle; a handle
 * owns one GPU (1 rank <-> 1 GPU), all device memory, its streams and its NCCL communicator.
 * Host arrays stay caller-owned and are only touched during upload/download calls.
 * There is no CPU fallback: without a CUDA device d3q19_create fails.
 *
 * Host array layouts are the reference's (column-major, var_inc.f90 / para.f90:418-503):
 *   f(0:18,lx,ly,lz)    f_aos[ip + 19*((ix-1) + lx*((iy-1) + ly*(iz-1)))]   ("canonical" = post-streaming,
 *                       exactly what the reference holds after collision_MRT returns)
 *   rho,ux,uy,uz,force  a[(ix-1) + lx*((iy-1) + ly*(iz-1))]
 *   ibnodes(0:lx+1,0:ly+1,0:lz+1) ghosted int32, -1 fluid / >0 solid (para.f90:442,447)
 * x is wall-normal and never decomposed; y is periodic and local; z is periodic and
 * slab-decomposed over ranks (rank r owns global planes [globalz, globalz+lz)).
 */
#ifndef D3Q19_B200_H
#define D3Q19_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D3Q19_NPOP 19
#define D3Q19_ABI_VERSION 1

/* Streaming scheme (how populations live in HBM). */
enum {
    D3Q19_SCHEME_AA = 0,   /* in-place, one array: even step local, odd step pull/push (default) */
    D3Q19_SCHEME_AB = 1,   /* two arrays, one-step pull                                          */
    D3Q19_SCHEME_AUTO = 2  /* AB when both arrays fit in < 45 % of free HBM, else AA             */
};

/* Arithmetic of the collision. */
enum {
    D3Q19_MATH_FAST = 0,   /* moment-space algebra with FMA contraction (production)              */
    D3Q19_MATH_STRICT = 1  /* the reference's expression order, no contraction: bit-exact vs oracle */
};

/* Where collision_MRT takes the conserved moments from. */
enum {
    D3Q19_MACRO_MAIN = 0,     /* rho,u = moments of f, in registers (main loop, main.f90:157)     */
    D3Q19_MACRO_PRERELAX = 1, /* rho = sum f (rhoupdat), u frozen in the device arrays (main.f90:70-90) */
    D3Q19_MACRO_EXTERNAL = 2  /* rho,u all read from the device arrays (set_macro / macrovar / avedensity) */
};

typedef struct d3q19_handle d3q19_handle;

/* All sizes are run-time (the reference fixes them at compile time, var_inc.f90:51). */
typedef struct d3q19_config {
    int32_t abi_version;      /* D3Q19_ABI_VERSION                                               */
    int32_t lx, ly, lz;       /* local extents; lx = nx, ly = ny (y not decomposed), lz = slab   */
    int32_t nx, ny, nz;       /* global extents                                                  */
    int32_t globalz;          /* first owned global z plane (0-based), para.f90:259-261          */
    int32_t rank, nranks;     /* z-slab ring; mzp/mzm = (rank +- 1) mod nranks, para.f90:266-267  */
    int32_t device;           /* CUDA ordinal (normally the local rank)                          */
    int32_t scheme;           /* D3Q19_SCHEME_*                                                  */
    int32_t math;             /* D3Q19_MATH_*                                                    */
    int32_t ipart;            /* para.f90:332 -- enables the solid-node paths                    */
    int32_t overlap;          /* 1: boundary/interior split with exchange on a second stream     */
    /* tuning knobs; 0 = the measured default.  (All behaviour switches live here or in d3q19_set_halo_mode: the
     * library reads no environment variable.)                                                      */
    int32_t nccl_max_ctas;    /* > 0: cap on the CTAs of NCCL's face send/recv (ncclConfig_t.maxCTAs); default NCCL's own
                                 choice -- every cap measured was a loss (profiles/r02e_two_gpus.md)                */
    int32_t pf_blocks;        /* in-place steps: L2 software-prefetch distance in thread blocks; default 128, < 0 off */
    int32_t halo_timeout_s;   /* peer-memory halo: seconds a wait for a neighbour's flag may last; default 30      */
    int32_t halo_split_min;   /* D3Q19_HALO_FUSED: slabs at least this thick run boundary and interior as two
                                 launches; default 64                                                               */
    int32_t force_idx64;      /* testing: 64-bit in-slab indices even when a population has < 2^32 elements         */
    /* MRT constants, para.f90:106-143 */
    double s1, s2, s4, s9, s10, s13, s16;
    double omegepsl, omegepslj, omegxx;
    double rhopart;           /* var_inc.f90:65 */
    double reserved_d[5];
    /* 128-byte ncclUniqueId made by rank 0 (d3q19_nccl_unique_id) and broadcast by the caller
     * (MPI_BCAST in the Fortran shim, torch.distributed in the Python host); ignored if nranks==1 */
    unsigned char nccl_id[128];
} d3q19_config;

/* ---- life cycle ---------------------------------------------------------------------- */
int d3q19_create(const d3q19_config *cfg, d3q19_handle **out);
int d3q19_destroy(d3q19_handle *h);
int d3q19_sync(d3q19_handle *h);
const char *d3q19_last_error(void);
int d3q19_nccl_unique_id(unsigned char out[128]);
int d3q19_device_count(int32_t *n);

/* ---- halo in NVLink peer memory (optional, nranks > 1) ----------------------------------------
 * Replaces the MPI_ISEND/IRECV/WAITALL ghost exchange (collision.f90:349-356) inside the step:
 * the boundary planes store their five outgoing populations straight into the neighbour GPU's
 * array and raise a flag there.  Bootstrap: every rank exports a blob of
 * D3Q19_IPC_BYTES, the caller all-gathers them (MPI_ALLGATHER in the Fortran shim,
 * torch.distributed in Python) and hands the nranks blobs, in rank order, to d3q19_ipc_connect.
 * Collective: every rank must call both.  Without it the exchange goes through NCCL send/recv. */
#define D3Q19_IPC_BYTES 256
int d3q19_ipc_export(d3q19_handle *h, unsigned char *blob);
int d3q19_ipc_connect(d3q19_handle *h, const unsigned char *blobs_in_rank_order);
/* How the peer-memory halo moves the faces (call before the first step):
 *   D3Q19_HALO_FUSED  the boundary planes of the step kernel store into the neighbour's array themselves
 *   D3Q19_HALO_PUT    plain step kernels; the COPY ENGINES move each crossing population from my array into the
 *                     neighbour's (ten device-to-device copies over NVLink per step, no SM involved) while the
 *                     interior is computed; a one-thread kernel then raises the neighbours' flags              */
enum { D3Q19_HALO_FUSED = 0, D3Q19_HALO_PUT = 1 };
int d3q19_set_halo_mode(d3q19_handle *h, int32_t mode);

/* ---- state transfer (canonical AoS layout, see above) --------------------------------- */
int d3q19_upload_f(d3q19_handle *h, const double *f_aos);
int d3q19_download_f(d3q19_handle *h, double *f_aos);
/* any pointer may be NULL (left untouched) */
int d3q19_set_macro(d3q19_handle *h, const double *rho, const double *ux, const double *uy, const double *uz);
int d3q19_download_macro(d3q19_handle *h, double *rho, double *ux, double *uy, double *uz);

/* initvel + initpop evaluated on the device (initial.f90:75-147 with ivel, :19-46): log-law mean
 * profile, the A9 perturbation block, plus a counter-based uniform noise of amplitude noise_amp
 * (splitmix64 of seed/component/global node -- SURVEY.md 8(d) synthetic inputs).  For fields too
 * large to stage through the host (configs[3], 150 GB per GPU).  No halo exchange is needed:
 * the velocity is a pure function of the global node coordinates.                          */
int d3q19_init_channel(d3q19_handle *h, double ustar, double ystar, double A9, double noise_amp,
                       uint64_t seed, int32_t ivel);

/* ---- forcing (FORCING / FORCINGP) ------------------------------------------------------ */
int d3q19_set_force_uniform(d3q19_handle *h, double fx, double fy, double fz);
int d3q19_set_force_field(d3q19_handle *h, const double *fx, const double *fy, const double *fz);
/* FORCINGP (collision.f90:529-602) evaluated on the device for step `istep`: the uniform force plus the
 * sinusoidal perturbation in two near-wall x-bands, written straight into the device force field. */
int d3q19_forcingp(d3q19_handle *h, int32_t istep, double force_in_y);
int d3q19_download_force_field(d3q19_handle *h, double *fx, double *fy, double *fz);

/* ---- the time step --------------------------------------------------------------------- */
/* one collision_MRT (collide + force + propagate + wall bounce-back + ghost exchange) */
int d3q19_collide_stream(d3q19_handle *h, int32_t macro_mode);
/* nsteps back-to-back main-loop steps (MACRO_MAIN), enqueued without host synchronisation */
int d3q19_run(d3q19_handle *h, int32_t nsteps);
/* macrovar on the device arrays (fluid branch + solid branch when ipart) */
int d3q19_macrovar(d3q19_handle *h);
int d3q19_rhoupdat(d3q19_handle *h);
/* global mean over fluid nodes removed from the device rho array; the next collide_stream
 * sees the shifted density exactly as the reference's does (collision.f90:505-511)       */
int d3q19_avedensity(d3q19_handle *h, double *rhomean, int64_t *nfluid_total);
/* rho,ux,uy,uz of one local node (1-based ix,iy,iz) computed from the current populations:
 * the reference's `probe` regression vector (saveload.f90:4059-4100)                      */
int d3q19_probe(d3q19_handle *h, int32_t ix, int32_t iy, int32_t iz, double out4[4]);
/* device-side pre-relaxation loop (main.f90:70-90): repeats rhoupdat + collision_MRT with u
 * frozen until max|rho - rho_prev| <= tol (all ranks) or iters > maxiter                  */
int d3q19_prerelax(d3q19_handle *h, double tol, int32_t maxiter, int32_t *iters, double *rhoerrmax);

/* ---- particles (ibnodes / isnodes, var_inc.f90:113-114) -------------------------------- */
int d3q19_set_solid_mask(d3q19_handle *h, const int32_t *ibnodes_ghosted, const int32_t *isnodes);
int d3q19_set_particles(d3q19_handle *h, int32_t npart, const double *ypglb, const double *wp, const double *omgp);

/* Device-side particle bookkeeping.  The reference snapshot does not contain its particle library
 * (main.f90:64 names partlib.f90), so these follow the published algorithms its
 * data structures point to and are "parity unpinned" (DESIGN.md section 7).  Entry points carry the
 * names of the reference's per-phase timers (var_inc.f90:166-168).                           */
typedef struct d3q19_particle_params {
    double rad;                 /* var_inc.f90:67                                            */
    double rho0;                /* var_inc.f90:65: density in the moving-wall / force terms  */
    double mingap, mingap_w;    /* var_inc.f90:67                                            */
    double stf0, stf1, stf0_w, stf1_w;   /* para.f90:356-360                                 */
    double fscale;              /* force scale of the repulsion (0 = off)                    */
    double gforce[3];           /* constant body force on each particle (gravity - buoyancy) */
    int64_t maxlink;            /* link capacity of ALL particles (split evenly into per-particle segments);
                                   0 = 8 npart 4 pi (rad+1)^2 (cf. para.f90:371); an overflow is reported by d3q19_sync */
} d3q19_particle_params;
int d3q19_particles_init(d3q19_handle *h, int32_t npart, const d3q19_particle_params *prm);
/* solid mask (ibnodes/isnodes, ghost planes included) and boundary links from the particle table.  A no-op while the
   table has not changed since the last build (d3q19_set_particles / d3q19_beads_move invalidate); every entry point that
   reads the mask (collide_stream, macrovar, avedensity, diag, profiles) builds it first if it is stale */
int d3q19_beads_links(d3q19_handle *h, int64_t *nlink_local);
/* interpolated bounce-back on every link after a collide_stream; force and torque reduced over ranks */
int d3q19_beads_collision(d3q19_handle *h);
int d3q19_beads_lubforce(d3q19_handle *h);
int d3q19_beads_move(d3q19_handle *h);
/* refill of the nodes uncovered by the last move (after d3q19_beads_links rebuilt the mask).  With nranks > 1 the call
   is COLLECTIVE: every rank sends the 19 canonical populations of its two boundary planes to its neighbours first, so
   that a refill next to a slab face uses the same source nodes as on a single domain                              */
int d3q19_beads_filling(d3q19_handle *h, int64_t *nfilled_local);
/* links (if stale); collide_stream; beads_collision; [lubforce; move; links; filling]         */
int d3q19_particle_step(d3q19_handle *h, int32_t move);
int d3q19_get_particles(d3q19_handle *h, double *ypglb, double *wp, double *omgp, double *fHIp, double *torqp);
/* the boundary links of this slab: global 1-based fluid-node coordinates, direction into the solid, particle (1-based),
 * fraction q of the link on the fluid side.  The list is a SET: the particles' segments come one after the other, inside a
 * segment the order depends on the run -- sort (particle, z, y, x, direction) before comparing two lists.            */
int d3q19_get_links(d3q19_handle *h, int64_t capacity, int32_t *x, int32_t *y, int32_t *z, int32_t *ip,
                    int32_t *part, double *q, int64_t *nlink);
int d3q19_get_mask(d3q19_handle *h, int32_t *own_lx_ly_lz);

/* ---- device-side statistics ------------------------------------ */
/* per-x-plane sums over the local (y,z) of ux,uy,uz,ux^2,uy^2,uz^2,uxuy,uxuz,uyuz,rho,rho^2
 * (statistc, saveload.f90:1241-1300), all-reduced over ranks; out is [11][lx]             */
int d3q19_profiles(d3q19_handle *h, double *out_11_by_lx);
/* statistc2 (saveload.f90:1348-1502): with a solid mask the sums run over the fluid nodes only; row 11 is
 * the number of fluid nodes of each x-plane (nfluid, :1428), exact in fp64; out is [12][lx]            */
int d3q19_profiles2(d3q19_handle *h, double *out_12_by_lx);

/* diag (saveload.f90:1507-1676) from the current populations, reduced over ranks.  out[14]: vmax,
 * imout, jmout, kmout (global, 1-based, first occurrence in the reference's loop order), umean, vmean,
 * wmean, urms, vrms, wrms (divided by ustar like the reference), volf, rhomax, rhomin, nfluid.       */
int d3q19_diag(d3q19_handle *h, double ustar, double *out14);

/* vortcalc + exchng8 (saveload.f90:3929-4054): vorticity of the device velocity
 * field (call d3q19_macrovar first) by central differences, one-sided at the walls, z-neighbour planes
 * exchanged over NCCL, solid nodes = twice the particle's angular velocity.  Bit-identical to the
 * reference's expression order.  The arrays are the reference's ox,oy,oz(lx,ly,lz) (var_inc.f90:140).  */
int d3q19_vortcalc(d3q19_handle *h);
int d3q19_download_vort(d3q19_handle *h, double *ox, double *oy, double *oz);

/* Local strain rate from the non-equilibrium moments (first loop nest of sijstat00, saveload.f90:2031-2091;
 * SURVEY.md 8(f) rank 4): Sij*Sij of every fluid node from its populations and the device rho,u (call d3q19_macrovar
 * first) -- node-local, reuses collision_MRT's sums, bit-identical to the reference's expression order; solid
 * nodes get 0.  The local dissipation rate is 2 visc Sij Sij (:1987-1989).  Array layout (lx,ly,lz).            */
int d3q19_sijstat(d3q19_handle *h);
int d3q19_download_sij2(d3q19_handle *h, double *sij2);

/* ---- measurement ----------------------------------------------------------------------- */
/* CUDA events on the stream the step kernels are launched on */
int d3q19_timer_start(d3q19_handle *h);
int d3q19_timer_stop(d3q19_handle *h, float *elapsed_ms);
/* counters: [0] step kernels launched, [1] other kernels launched, [2] NCCL ops,
 *           [3] steps taken, [4] bytes of populations resident, [5] storage phase,
 *           [6] x pitch, [7] scheme in use                                                 */
int d3q19_get_counters(d3q19_handle *h, int64_t out[8]);
/* per-step timeline (development aid, tools/step_timeline.py): four CUDA timing events per step for the next
 * max_steps steps -- before / after the boundary-plane launch, after the interior launch (compute stream), after
 * the z-face exchange (exchange stream); fetch returns milliseconds relative to the first mark, out[4*s + k]   */
int d3q19_trace_enable(d3q19_handle *h, int32_t max_steps);
int d3q19_trace_fetch(d3q19_handle *h, int32_t *nsteps, float *out, int32_t capacity_steps);

/* ---- the shim state machine (what the replacement collision.f90 calls) ------------------ */
/* These carry the download policy of SURVEY.md section 8(b): device-resident f is
 * authoritative; host rho,u are refreshed only on steps where the intact driver reads them. */
typedef struct d3q19_shim_arrays {
    double *f;                       /* f(0:18,lx,ly,lz)                       */
    double *rho, *ux, *uy, *uz;      /* (lx,ly,lz)                             */
    double *force_realx, *force_realy, *force_realz;
    int32_t *ibnodes, *isnodes;
    int32_t ndiag, nflowout, nsteps_total, istep0;   /* var_inc.f90:58-59, para.f90:43-45 */
    int32_t ntime;                   /* wall-clock guard cadence, var_inc.f90:59 / main.f90:197 (0 = never) */
    int32_t prerelax_maxiter;        /* main.f90:85 (15000); 0 = library default 15000     */
    double rhoepsl;                  /* para.f90:285; pre-relaxation exit tolerance        */
} d3q19_shim_arrays;

int d3q19_shim_bind(d3q19_handle *h, const d3q19_shim_arrays *a);
/* the same with the arrays as by-reference arguments (a Fortran caller needs no C_LOC/TARGET) */
int d3q19_shim_bind_arrays(d3q19_handle *h, double *f, double *rho, double *ux, double *uy, double *uz,
                           double *force_realx, double *force_realy, double *force_realz, int32_t *ibnodes,
                           int32_t *isnodes, int32_t has_isnodes, int32_t ndiag, int32_t nflowout,
                           int32_t nsteps_total, int32_t istep0, int32_t ntime, int32_t prerelax_maxiter,
                           double rhoepsl);
/* change the output cadence / loop bounds the download policy keys on (main.f90:142,171,184) */
int d3q19_shim_set_schedule(d3q19_handle *h, int32_t ndiag, int32_t nflowout, int32_t nsteps_total, int32_t istep0);
int d3q19_shim_forcing(d3q19_handle *h, double force_in_y, double force_mag);
/* rhoupdat also forms max|rho_new - rho_old| over all ranks on the device: it is the number
 * main.f90:79-80 is about to compute, so the shim knows when the intact driver will leave the
 * pre-relaxation loop (main.f90:85) and makes the host f current for saveinitflow (main.f90:101)
 * in the collision_MRT of that very iteration -- and in no other.                          */
int d3q19_shim_rhoupdat(d3q19_handle *h);
int d3q19_shim_collision_mrt(d3q19_handle *h);
/* last pre-relaxation error / iteration index seen by d3q19_shim_rhoupdat */
int d3q19_shim_prerelax_state(d3q19_handle *h, double *rhoerrmax, int32_t *iteration);
int d3q19_shim_macrovar(d3q19_handle *h, int32_t istep);
int d3q19_shim_avedensity(d3q19_handle *h);
/* make the host copy of f current (before savecntdflow / saveinitflow, saveload.f90:120,227) */
int d3q19_shim_sync_f_to_host(d3q19_handle *h);
/* the host copy of f changed (after loadcntdflow / initpop) */
int d3q19_shim_sync_f_to_device(d3q19_handle *h);

#ifdef __cplusplus
}
#endif
#endif
