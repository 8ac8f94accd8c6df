/*
 * d3q19_oracle.h -- CPU restatement of the UDel-CFD D3Q19 Channel-Flow time-step path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker (or as the reported CPU baseline), never as the thing shipped.
 *
 * PARITY PINNING: pinned to the reference itself for the fluid path.  The reference is Fortran 90
 * + MPI and this image has no Fortran compiler and no MPI, so it cannot be built as shipped;
 * instead oracle/f90toc.py machine-translates the reference's own collision.f90 / para.f90 /
 * initial.f90 / var_inc.f90, from the sources where they lie under /root/reference, into C
 * (oracle/_ref/, git-ignored) and runs it with one thread per MPI rank (oracle/ref_runtime.h).
 * This restatement agrees with that translation BIT FOR BIT on every case of
 * tests/test_oracle_ref.py (1x1 ... 3x2 rank grids, uneven splits, all three MRT types,
 * pre-relaxation, avedensity, force fields, solid nodes), and reproduces bit for bit the
 * golden vectors the translation generated (tests/golden/*.npz, tests/test_golden.py) --
 * those travel to the GPU box, where /root/reference does not exist.  Further anchors:
 *   (1) the analytic start-up / steady Poiseuille known-answers the reference itself embeds
 *       in saveload.f90:921-935 (tests/test_oracle.py),
 *   (2) an independent matrix-form ("textbook") numpy restatement (oracle/textbook.py).
 * The reference ships no tests, fixtures or golden vectors of its own.
 * The PARTICLE path (interpolated bounce-back, refill, force) is "parity unpinned": the
 * reference snapshot does not contain partlib.f90.
 *
 * Every function cites the reference file:line (relative to Channel-Flow/) it follows.
 * Arrays use the Fortran column-major layouts of var_inc.f90 / para.f90:418-503:
 *   f(0:18,lx,ly,lz)            -> f[ip + 19*((ix-1) + lx*((iy-1) + ly*(iz-1)))]
 *   rho,ux,uy,uz,force(lx,ly,lz)-> a[(ix-1) + lx*((iy-1) + ly*(iz-1))]
 *   ibnodes(0:lx+1,0:ly+1,0:lz+1)-> ib[ix + (lx+2)*(iy + (ly+2)*iz)]
 * Build with -ffp-contract=off so that every expression is the IEEE evaluation of the
 * Fortran source order.
 */
#ifndef D3Q19_ORACLE_H
#define D3Q19_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NPOP 19

/* Scalars of `module var_inc` that the path reads (var_inc.f90:51-99) and that
 * `para` sets (para.f90:59-214). */
typedef struct orc_para {
    int nx, ny, nz;            /* global sizes; lx = nx (x is never decomposed)          */
    int nprocY, nprocZ;        /* rank grid, para.f90:219-220                             */
    int laminar;               /* para.f90:59 laminarFlow                                 */
    int MRTtype;               /* para.f90:70,86                                          */
    int ivel;                  /* para.f90:69,85                                          */
    double visc, Rstar, ustar, force_in_y, ystar, force_mag, rho0;
    double tau, s1, s2, s4, s9, s10, s13, s16;
    double omegepsl, omegepslj, omegxx;
    double coef1, coef2, coef3, coef4, coef5, coef3i, coef4i;
    double val1, val2, val3, val4, val5, val6, val7, val8, val9;
    double val1i, val2i, val3i, val4i, val5i, val6i, val7i, val8i, val9i;
    double ww0, ww1, ww2;
    double pi, pi2;
    double rhopart;
    int ipart;
    int cix[ORC_NPOP], ciy[ORC_NPOP], ciz[ORC_NPOP], ipopp[ORC_NPOP];
    int ipswap[9], ipstay[10];
} orc_para;

/* One MPI rank's share of `module var_inc` (allocarray, para.f90:418-503). */
typedef struct orc_rank {
    int myid, indy, indz;
    int lx, ly, lz;
    int globaly, globalz;
    int myp, mym, mzp, mzm;
    double *f;
    double *rho, *ux, *uy, *uz;
    double *force_realx, *force_realy, *force_realz;
    int32_t *ibnodes;          /* ghosted, -1 fluid / >0 solid                            */
    int32_t *isnodes;          /* (lx,ly,lz) owning particle id, only if ipart            */
    /* collision_MRT's automatic arrays (collision.f90:31-32), 19 slots each              */
    double *tmpymS, *tmpypS;   /* (0:18, lx, 0:lz+1)                                      */
    double *tmpzmS, *tmpzpS;   /* (0:18, lx, ly)                                          */
    /* collisionExchnge's 5-slot buffers (collision.f90:290-291)                          */
    double *ymS5, *ypS5, *ymR5, *ypR5;   /* (5, lx, 0:lz+1)                               */
    double *zmS5, *zpS5, *zmR5, *zpR5;   /* (5, lx, ly)                                   */
} orc_rank;

typedef struct orc_world {
    orc_para p;
    int nproc;
    orc_rank *r;
    /* particle tables (global sized, para.f90:454-460), only for macrovar's solid branch */
    int npart;
    double *ypglb, *wp, *omgp;  /* (3,npart) */
} orc_world;

/* para.f90:59-214 -- parameter sets, MRT constants, lattice tables. */
void orc_para_init(orc_para *p, int nx, int ny, int nz, int laminar, int nprocY, int nprocZ);
/* Re-derive the MRT relaxation set after changing p->MRTtype / p->visc (para.f90:106-143). */
void orc_para_set_mrt(orc_para *p);

/* para.f90:229-276 + allocarray: build all ranks of an nprocY x nprocZ topology. */
orc_world *orc_world_create(const orc_para *p);
void orc_world_destroy(orc_world *w);

/* initial.f90:75-147 (A9 is the reference's hard-wired 0.0 unless overridden). */
void orc_initvel(orc_world *w, double A9);
/* initial.f90:19-46 */
void orc_initpop(orc_world *w);
/* collision.f90:515-527 */
void orc_forcing(orc_world *w);
/* collision.f90:24-268: the local sweep of one rank (fills the tmp?S send buffers). */
void orc_collision_local(const orc_para *p, orc_rank *r);
/* collision.f90:281-372 split at the two MPI_WAITALLs so that a test may move the
 * buffers itself (gloo) or let orc_exchange do it in-process. */
void orc_pack_y(const orc_para *p, orc_rank *r);
void orc_unpack_y_pack_z(const orc_para *p, orc_rank *r);
void orc_unpack_z(const orc_para *p, orc_rank *r);
/* In-process stand-in for the MPI_ISEND/IRECV/WAITALL pairs (collision.f90:309-314,351-356). */
void orc_deliver_y(orc_world *w);
void orc_deliver_z(orc_world *w);
/* collision_MRT for every rank = local sweeps + exchange (threads over ranks if OpenMP). */
void orc_collision_MRT(orc_world *w);
/* collision.f90:378-463 */
void orc_macrovar(orc_world *w);
/* collision.f90:469-480 */
void orc_rhoupdat(orc_world *w);
/* saveload.f90:3929-4054 (vortcalc + exchng8): vorticity of ux,uy,uz into global (lx,ny,nz) arrays. */
void orc_vortcalc(const orc_world *w, double *ox_global, double *oy_global, double *oz_global);
/* saveload.f90:2031-2091 (first loop nest of sijstat00): Sij*Sij of the fluid nodes into a global (nx,ny,nz) array. */
void orc_sijstat(const orc_world *w, double *sij2_global);
/* collision.f90:487-513; returns rhomean, writes the global fluid-node count. */
double orc_avedensity(orc_world *w, int64_t *nfluidtotal);

/* Gather / scatter between the rank-local arrays and one global array in the
 * same Fortran layout with (lx,ny,nz) extents (test convenience, not in the reference). */
void orc_gather_f(const orc_world *w, double *fglobal);
void orc_scatter_f(orc_world *w, const double *fglobal);
void orc_gather_scalar(const orc_world *w, int which, double *aglobal); /* 0 rho 1 ux 2 uy 3 uz */
void orc_scatter_scalar(orc_world *w, int which, const double *aglobal);
void orc_scatter_ibnodes(orc_world *w, const int32_t *ib_global_noghost, const int32_t *is_global);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
