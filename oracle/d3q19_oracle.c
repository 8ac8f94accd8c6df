/*
 * d3q19_oracle.c -- CPU restatement of the reference's time-step path.  See d3q19_oracle.h
 * for the "test infrastructure only" and "parity unpinned" statements.
 *
 * The statements below follow the Fortran source order so that, compiled with
 * -ffp-contract=off, each expression is the IEEE-754 double evaluation of the reference
 * text (left-to-right for equal precedence; `-r8` makes every real and literal 8 bytes,
 * Makefile:29).  Line numbers cite /root/reference/Channel-Flow/.
 */
#include "d3q19_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define NPOP ORC_NPOP

/* ---- Fortran index maps ------------------------------------------------------------ */
#define F_(r, ip, ix, iy, iz) \
    ((r)->f[(size_t)(ip) + NPOP * ((size_t)((ix)-1) + (size_t)(r)->lx * ((size_t)((iy)-1) + (size_t)(r)->ly * (size_t)((iz)-1)))])
#define S_(r, a, ix, iy, iz) \
    ((a)[(size_t)((ix)-1) + (size_t)(r)->lx * ((size_t)((iy)-1) + (size_t)(r)->ly * (size_t)((iz)-1))])
#define IB_(r, ix, iy, iz) \
    ((r)->ibnodes[(size_t)(ix) + (size_t)((r)->lx + 2) * ((size_t)(iy) + (size_t)((r)->ly + 2) * (size_t)(iz))])
/* tmpy?S(0:18, lx, 0:lz+1) and tmpz?S(0:18, lx, ly)  (collision.f90:31-32) */
#define TY19(r, a, ip, ix, k) ((a)[(size_t)(ip) + NPOP * ((size_t)((ix)-1) + (size_t)(r)->lx * (size_t)(k))])
#define TZ19(r, a, ip, ix, j) ((a)[(size_t)(ip) + NPOP * ((size_t)((ix)-1) + (size_t)(r)->lx * (size_t)((j)-1))])
/* 5-slot buffers (1:5, lx, 0:lz+1) and (1:5, lx, ly)  (collision.f90:290-291) */
#define TY5(r, a, s, ix, k) ((a)[(size_t)((s)-1) + 5 * ((size_t)((ix)-1) + (size_t)(r)->lx * (size_t)(k))])
#define TZ5(r, a, s, ix, j) ((a)[(size_t)((s)-1) + 5 * ((size_t)((ix)-1) + (size_t)(r)->lx * (size_t)((j)-1))])

/* ---- a tiny pthread "parallel for over ranks" (stands in for one process per MPI rank) */
static int g_threads = 0;

int orc_num_threads(void)
{
    if (g_threads <= 0) {
        const char *e = getenv("ORC_THREADS");
        long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
        g_threads = n > 0 ? (int)n : 1;
    }
    return g_threads;
}

void orc_set_num_threads(int n) { g_threads = n > 0 ? n : 1; }

typedef struct orc_pf {
    int n, next;
    void (*fn)(void *ctx, int id);
    void *ctx;
} orc_pf;

static void *orc_pf_worker(void *arg)
{
    orc_pf *pf = (orc_pf *)arg;
    for (;;) {
        int id = __atomic_fetch_add(&pf->next, 1, __ATOMIC_RELAXED);
        if (id >= pf->n) break;
        pf->fn(pf->ctx, id);
    }
    return 0;
}

static void orc_parallel_for(int n, void (*fn)(void *, int), void *ctx)
{
    int nt = orc_num_threads(), t;
    orc_pf pf;
    pthread_t th[256];
    pf.n = n; pf.next = 0; pf.fn = fn; pf.ctx = ctx;
    if (nt > n) nt = n;
    if (nt > 256) nt = 256;
    if (nt <= 1) { orc_pf_worker(&pf); return; }
    for (t = 1; t < nt; ++t) pthread_create(&th[t], 0, orc_pf_worker, &pf);
    orc_pf_worker(&pf);
    for (t = 1; t < nt; ++t) pthread_join(th[t], 0);
}

/* ---- para.f90:106-143 ---------------------------------------------------------------- */
void orc_para_set_mrt(orc_para *p)
{
    p->tau = 3.0 * p->visc + 0.5;                 /* para.f90:106 */
    p->s9 = 1.0 / p->tau;                         /* :107 */
    p->s13 = p->s9;                               /* :108 */
    switch (p->MRTtype) {
    case 1:                                       /* :111-120 */
        p->s1 = 1.5; p->s2 = 1.4; p->s4 = 1.2; p->s10 = 1.4; p->s16 = 1.98;
        p->omegepsl = 0.0; p->omegepslj = -475.0 / 63.0; p->omegxx = 0.0;
        break;
    case 2:                                       /* :121-130 */
        p->s1 = p->s9; p->s2 = p->s9; p->s4 = p->s9; p->s10 = p->s9; p->s16 = p->s9;
        p->omegepsl = 3.0; p->omegepslj = -11.0 / 2.0; p->omegxx = -1.0 / 2.0;
        break;
    default:                                      /* case(3) :131-141 */
        p->s1 = 1.8; p->s2 = p->s1; p->s4 = p->s9; p->s10 = p->s1; p->s16 = p->s1;
        p->omegepsl = 3.0; p->omegepslj = -11.0 / 2.0; p->omegxx = -1.0 / 2.0;
        break;
    }
}

/* ---- para.f90:59-214 ----------------------------------------------------------------- */
void orc_para_init(orc_para *p, int nx, int ny, int nz, int laminar, int nprocY, int nprocZ)
{
    static const int cix[NPOP] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
    static const int ciy[NPOP] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
    static const int ciz[NPOP] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
    static const int ipopp[NPOP] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
    static const int ipswap[9] = {2, 4, 6, 9, 10, 13, 14, 17, 18};
    static const int ipstay[10] = {0, 1, 3, 5, 7, 8, 11, 12, 15, 16};

    memset(p, 0, sizeof(*p));
    p->nx = nx; p->ny = ny; p->nz = nz;
    p->nprocY = nprocY; p->nprocZ = nprocZ;
    p->laminar = laminar;
    p->rho0 = 1.0; p->rhopart = 1.0;              /* var_inc.f90:65 */
    p->pi = 4.0 * atan(1.0);                      /* var_inc.f90:71 */
    p->pi2 = 2.0 * p->pi;                         /* var_inc.f90:72 */
    p->ipart = 0;                                 /* para.f90:332 */
    if (!laminar) {                               /* para.f90:61-70 */
        p->visc = 0.0036;
        p->Rstar = 180.0;
        p->ustar = 2.0 * p->Rstar * p->visc / (double)nx;
        p->force_in_y = 2. * p->rho0 * p->ustar * p->ustar / (double)nx;
        p->ystar = p->visc / p->ustar;
        p->force_mag = 1.0;
        p->ivel = 1;
        p->MRTtype = 1;
    } else {                                      /* para.f90:75-88 */
        p->Rstar = 20;
        p->ustar = 0.05;
        p->visc = 2.0 * p->ustar * (double)nx / p->Rstar;
        p->force_in_y = 8.0 * p->visc * p->ustar / ((double)nx * (double)nx);
        p->ystar = p->visc / p->ustar;
        p->force_mag = 1.0;
        p->ivel = 0;
        p->MRTtype = 2;
    }
    orc_para_set_mrt(p);

    p->coef1 = -2.0 / 3.0;                        /* para.f90:143-150 */
    p->coef2 = -11.0; p->coef3 = 8.0; p->coef4 = -4.0; p->coef5 = 2.0;
    p->coef3i = 1.0 / p->coef3; p->coef4i = 1.0 / p->coef4;
    p->val1 = 19.0; p->val2 = 2394.0; p->val3 = 252.0; p->val4 = 10.0; p->val5 = 40.0;
    p->val6 = 36.0; p->val7 = 72.0; p->val8 = 12.0; p->val9 = 24.0;
    p->val1i = 1.0 / p->val1; p->val2i = 1.0 / p->val2; p->val3i = 1.0 / p->val3;
    p->val4i = 1.0 / p->val4; p->val5i = 1.0 / p->val5; p->val6i = 1.0 / p->val6;
    p->val7i = 1.0 / p->val7; p->val8i = 1.0 / p->val8; p->val9i = 1.0 / p->val9;
    p->ww0 = 1.0 / 3.0; p->ww1 = 1.0 / 18.0; p->ww2 = 1.0 / 36.0;   /* :172-174 */
    memcpy(p->cix, cix, sizeof cix); memcpy(p->ciy, ciy, sizeof ciy);      /* :201-203 */
    memcpy(p->ciz, ciz, sizeof ciz); memcpy(p->ipopp, ipopp, sizeof ipopp);/* :206 */
    memcpy(p->ipswap, ipswap, sizeof ipswap);                               /* :209 */
    memcpy(p->ipstay, ipstay, sizeof ipstay);                               /* :210 */
}

/* ---- para.f90:229-276 + allocarray (para.f90:418-503) ------------------------------- */
static void *xcalloc(size_t n, size_t sz)
{
    void *q = calloc(n ? n : 1, sz);
    if (!q) abort();
    return q;
}

orc_world *orc_world_create(const orc_para *p)
{
    orc_world *w = (orc_world *)xcalloc(1, sizeof(*w));
    const int nprocY = p->nprocY, nprocZ = p->nprocZ, nproc = nprocY * nprocZ;
    const int ny = p->ny, nz = p->nz;
    int *mpily, *mpilz;
    int id, i;
    w->p = *p;
    w->nproc = nproc;
    w->r = (orc_rank *)xcalloc((size_t)nproc, sizeof(orc_rank));
    mpily = (int *)xcalloc((size_t)nproc, sizeof(int));
    mpilz = (int *)xcalloc((size_t)nproc, sizeof(int));
    for (id = 0; id < nproc; ++id) {
        orc_rank *r = &w->r[id];
        r->myid = id;
        r->indy = id % nprocY;                                   /* :229 */
        r->indz = id / nprocY;                                   /* :230 */
        r->lx = p->nx;
        if (r->indy < ny - nprocY * (ny / nprocY))               /* :233-237 */
            r->ly = (ny - ny % nprocY) / nprocY + 1;
        else
            r->ly = (ny - ny % nprocY) / nprocY;
        if (r->indz < nz - nprocZ * (nz / nprocZ))               /* :240-244 */
            r->lz = (nz - nz % nprocZ) / nprocZ + 1;
        else
            r->lz = (nz - nz % nprocZ) / nprocZ;
        mpily[id] = r->ly;                                       /* :250-251 Allgather */
        mpilz[id] = r->lz;
    }
    for (id = 0; id < nproc; ++id) {
        orc_rank *r = &w->r[id];
        const int indy = r->indy, indz = r->indz;
        size_t n3, nyb, nzb;
        r->globaly = 0; r->globalz = 0;                          /* :254-261 */
        for (i = 0; i <= indy - 1; ++i) r->globaly += mpily[indz * nprocY + i];
        for (i = 0; i <= indz - 1; ++i) r->globalz += mpilz[i * nprocY + indy];
        r->mzp = ((indz + 1) % nprocZ) * nprocY + indy;          /* :266 */
        r->mzm = ((indz + nprocZ - 1) % nprocZ) * nprocY + indy; /* :267 */
        r->myp = indz * nprocY + (indy + 1) % nprocY;            /* :269 */
        r->mym = indz * nprocY + (indy + nprocY - 1) % nprocY;   /* :270 */

        n3 = (size_t)r->lx * r->ly * r->lz;
        r->f = (double *)xcalloc(n3 * NPOP, sizeof(double));
        r->rho = (double *)xcalloc(n3, sizeof(double));
        r->ux = (double *)xcalloc(n3, sizeof(double));
        r->uy = (double *)xcalloc(n3, sizeof(double));
        r->uz = (double *)xcalloc(n3, sizeof(double));
        r->force_realx = (double *)xcalloc(n3, sizeof(double));
        r->force_realy = (double *)xcalloc(n3, sizeof(double));
        r->force_realz = (double *)xcalloc(n3, sizeof(double));
        {
            size_t nib = (size_t)(r->lx + 2) * (r->ly + 2) * (r->lz + 2), q;
            r->ibnodes = (int32_t *)xcalloc(nib, sizeof(int32_t));
            for (q = 0; q < nib; ++q) r->ibnodes[q] = -1;        /* para.f90:447 */
            r->isnodes = (int32_t *)xcalloc(n3, sizeof(int32_t));
            for (q = 0; q < n3; ++q) r->isnodes[q] = -1;         /* para.f90:499 */
        }
        nyb = (size_t)r->lx * (r->lz + 2);
        nzb = (size_t)r->lx * r->ly;
        r->tmpymS = (double *)xcalloc(nyb * NPOP, sizeof(double));
        r->tmpypS = (double *)xcalloc(nyb * NPOP, sizeof(double));
        r->tmpzmS = (double *)xcalloc(nzb * NPOP, sizeof(double));
        r->tmpzpS = (double *)xcalloc(nzb * NPOP, sizeof(double));
        r->ymS5 = (double *)xcalloc(nyb * 5, sizeof(double));
        r->ypS5 = (double *)xcalloc(nyb * 5, sizeof(double));
        r->ymR5 = (double *)xcalloc(nyb * 5, sizeof(double));
        r->ypR5 = (double *)xcalloc(nyb * 5, sizeof(double));
        r->zmS5 = (double *)xcalloc(nzb * 5, sizeof(double));
        r->zpS5 = (double *)xcalloc(nzb * 5, sizeof(double));
        r->zmR5 = (double *)xcalloc(nzb * 5, sizeof(double));
        r->zpR5 = (double *)xcalloc(nzb * 5, sizeof(double));
    }
    free(mpily);
    free(mpilz);
    return w;
}

void orc_world_destroy(orc_world *w)
{
    int id;
    if (!w) return;
    for (id = 0; id < w->nproc; ++id) {
        orc_rank *r = &w->r[id];
        free(r->f); free(r->rho); free(r->ux); free(r->uy); free(r->uz);
        free(r->force_realx); free(r->force_realy); free(r->force_realz);
        free(r->ibnodes); free(r->isnodes);
        free(r->tmpymS); free(r->tmpypS); free(r->tmpzmS); free(r->tmpzpS);
        free(r->ymS5); free(r->ypS5); free(r->ymR5); free(r->ypR5);
        free(r->zmS5); free(r->zpS5); free(r->zmR5); free(r->zpR5);
    }
    free(w->r);
    free(w->ypglb); free(w->wp); free(w->omgp);
    free(w);
}

/* ---- initial.f90:75-147 -------------------------------------------------------------- */
void orc_initvel(orc_world *w, double A9)
{
    const orc_para *p = &w->p;
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    const int nxh = (nx + 1) / 2;        /* var_inc.f90:51,54: nxh = nx7/2 with nx = nx7-1 */
    const double alpha = 1.0, beta9 = 1.0, cc = 60.0;                 /* :85-91 */
    const double ccc1 = -(double)ny / p->pi2 / alpha / p->ystar * A9 * p->ustar / cc / cc; /* :92 */
    int id;
    for (id = 0; id < w->nproc; ++id) {
        orc_rank *r = &w->r[id];
        const size_t n3 = (size_t)r->lx * r->ly * r->lz;
        int i, j, k;
        memset(r->ux, 0, n3 * sizeof(double));                        /* :94-96 */
        memset(r->uy, 0, n3 * sizeof(double));
        memset(r->uz, 0, n3 * sizeof(double));
        if (!p->ivel) continue;                                       /* :99 */
        for (i = 1; i <= nxh; ++i) {                                  /* :104-115 */
            double yplus = ((double)i - 0.5) / p->ystar;
            double u9;
            if (yplus < 10.8) {
                u9 = yplus * p->ustar;
            } else {
                u9 = log(yplus) / 0.41 + 5.0;
                u9 = u9 * p->ustar;
            }
            for (k = 1; k <= r->lz; ++k)
                for (j = 1; j <= r->ly; ++j) {
                    S_(r, r->uy, i, j, k) = u9;
                    S_(r, r->uy, nx + 1 - i, j, k) = u9;
                }
        }
        for (k = 1; k <= r->lz; ++k) {                                /* :119-144 */
            int kk = k + r->indz * r->lz;
            double z9 = p->pi2 * ((double)kk - 0.5) / (double)nz;
            for (j = 1; j <= r->ly; ++j) {
                int jj = j + r->indy * r->ly;
                double y9 = p->pi2 * ((double)jj - 0.5) / (double)ny;
                for (i = 1; i <= nxh; ++i) {
                    double yplus = ((double)i - 0.5) / p->ystar;
                    double ccc9 = exp(-yplus / cc);
                    double u9 = ccc1 * yplus * ccc9 * sin(alpha * y9 + beta9 * z9);
                    double ccc10;
                    S_(r, r->uy, i, j, k) = S_(r, r->uy, i, j, k) + u9;
                    S_(r, r->uy, nx + 1 - i, j, k) = S_(r, r->uy, nx + 1 - i, j, k) + u9;
                    ccc10 = A9 * p->ustar * (1. - ccc9 - yplus / cc * ccc9);
                    u9 = ccc10 * cos(alpha * y9 + beta9 * z9);
                    S_(r, r->ux, i, j, k) = S_(r, r->ux, i, j, k) + u9;
                    S_(r, r->ux, nx + 1 - i, j, k) = S_(r, r->ux, nx + 1 - i, j, k) + u9;
                }
            }
        }
    }
}

/* ---- initial.f90:19-46 --------------------------------------------------------------- */
void orc_initpop(orc_world *w)
{
    const orc_para *p = &w->p;
    int id;
    for (id = 0; id < w->nproc; ++id) {
        orc_rank *r = &w->r[id];
        const size_t n3 = (size_t)r->lx * r->ly * r->lz;
        size_t n;
        int ip;
        for (n = 0; n < n3; ++n) {
            double ux = r->ux[n], uy = r->uy[n], uz = r->uz[n];
            double usqr = ux * ux + uy * uy + uz * uz;                 /* :26 */
            double rho, G;
            usqr = 1.5 * usqr;                                         /* :27 */
            r->rho[n] = 0.0;                                           /* :32 */
            rho = r->rho[n];
            r->f[0 + NPOP * n] = p->ww0 * (rho - usqr);                /* :34 */
            for (ip = 1; ip <= 6; ++ip) {                              /* :36-39 */
                G = (p->cix[ip] * ux + p->ciy[ip] * uy + p->ciz[ip] * uz);
                r->f[ip + NPOP * n] = p->ww1 * (rho + 3.0 * G + 4.5 * G * G - usqr);
            }
            for (ip = 7; ip <= NPOP - 1; ++ip) {                       /* :41-44 */
                G = (p->cix[ip] * ux + p->ciy[ip] * uy + p->ciz[ip] * uz);
                r->f[ip + NPOP * n] = p->ww2 * (rho + 3.0 * G + 4.5 * G * G - usqr);
            }
        }
    }
}

/* ---- collision.f90:515-527 ----------------------------------------------------------- */
void orc_forcing(orc_world *w)
{
    const orc_para *p = &w->p;
    int id;
    for (id = 0; id < w->nproc; ++id) {
        orc_rank *r = &w->r[id];
        const size_t n3 = (size_t)r->lx * r->ly * r->lz;
        size_t n;
        for (n = 0; n < n3; ++n) {
            r->force_realx[n] = 0.0;
            r->force_realy[n] = p->force_in_y * p->force_mag;
            r->force_realz[n] = 0.0;
        }
    }
}

/* ---- collision.f90:24-268 ------------------------------------------------------------ */
void orc_collision_local(const orc_para *p, orc_rank *r)
{
    const int lx = r->lx, ly = r->ly, lz = r->lz;
    const int *cix = p->cix, *ciy = p->ciy, *ciz = p->ciz, *ipopp = p->ipopp;
    const double s1 = p->s1, s2 = p->s2, s4 = p->s4, s9 = p->s9, s10 = p->s10, s13 = p->s13, s16 = p->s16;
    const double coef1 = p->coef1, coef2 = p->coef2, coef3 = p->coef3, coef4 = p->coef4, coef5 = p->coef5;
    const double coef3i = p->coef3i, coef4i = p->coef4i;
    const double val1i = p->val1i, val2i = p->val2i, val3i = p->val3i, val4i = p->val4i, val5i = p->val5i,
                 val6i = p->val6i, val7i = p->val7i, val8i = p->val8i, val9i = p->val9i, val8 = p->val8;
    const double ww1 = p->ww1, ww2 = p->ww2;
    const double omegepsl = p->omegepsl, omegepslj = p->omegepslj, omegxx = p->omegxx;
    /* f9 and Fbar persist across nodes exactly like the Fortran locals do: a solid node
     * (goto 111) re-uses the previous node's values in the swap loop (collision.f90:54,244). */
    double f9[NPOP], Fbar[NPOP];
    int ip, ipi, ix, iy, iz, imove, jmove, kmove;
    memset(f9, 0, sizeof f9);
    memset(Fbar, 0, sizeof Fbar);

    for (iz = 1; iz <= lz; ++iz)                                       /* :50-52 */
    for (iy = 1; iy <= ly; ++iy)
    for (ix = 1; ix <= lx; ++ix) {
        if (!(IB_(r, ix, iy, iz) > 0)) {                               /* :54 */
            double rho9 = S_(r, r->rho, ix, iy, iz);                   /* :56-62 */
            double ux9 = S_(r, r->ux, ix, iy, iz);
            double uy9 = S_(r, r->uy, ix, iy, iz);
            double uz9 = S_(r, r->uz, ix, iy, iz);
            double ux9s = ux9 * ux9;
            double uy9s = uy9 * uy9;
            double uz9s = uz9 * uz9;
            double fx9 = S_(r, r->force_realx, ix, iy, iz);            /* :65-68 */
            double fy9 = S_(r, r->force_realy, ix, iy, iz);
            double fz9 = S_(r, r->force_realz, ix, iy, iz);
            double G3 = ux9 * fx9 + uy9 * fy9 + uz9 * fz9;
            double G1, G2, t1;
            double eqm1, eqm2, eqm3, eqm4, eqm5, eqm6, eqm7, eqm8, eqm9, eqm10, eqm11, eqm12, eqm13, eqm14, eqm15;
            double sum1, sum2, sum3, sum4, sum5, sum6, sum7, sum8, sum9, sum10, sum11;
            double evlm1, evlm2, evlm3, evlm4, evlm5, evlm6, evlm7, evlm8, evlm9, evlm10, evlm11, evlm12,
                   evlm13, evlm14, evlm15;
            double eqmc1, eqmc2, eqmc3, eqmc4, eqmc5, eqmc6, eqmc7, eqmc8, eqmc9, eqmc10, eqmc11, eqmc12,
                   eqmc13, eqmc14, eqmc15;
            double tl1, tl2, tl3, tl4, tl5, tl6, tl7, tl8, tl9, tl10, tl11, tl12, tl13, tl14, tl15, tl16,
                   tl17, tl18, tl19, tl20, tl21;
            double suma, sumb, sumc, sumd, sume, sumf, sumg, sumh, sumi, sumk, sump, sum67, sum89, sum1011;

            Fbar[0] = -G3;                                             /* :70 */
            for (ip = 1; ip <= 6; ++ip) {                              /* :72-76 */
                G1 = cix[ip] * fx9 + ciy[ip] * fy9 + ciz[ip] * fz9;
                G2 = cix[ip] * ux9 + ciy[ip] * uy9 + ciz[ip] * uz9;
                Fbar[ip] = ww1 * (3. * G1 + 9. * G1 * G2 - 3. * G3);
            }
            for (ip = 7; ip <= NPOP - 1; ++ip) {                       /* :78-82 */
                G1 = cix[ip] * fx9 + ciy[ip] * fy9 + ciz[ip] * fz9;
                G2 = cix[ip] * ux9 + ciy[ip] * uy9 + ciz[ip] * uz9;
                Fbar[ip] = ww2 * (3. * G1 + 9. * G1 * G2 - 3. * G3);
            }
            for (ip = 0; ip < NPOP; ++ip)                              /* :84 */
                f9[ip] = F_(r, ip, ix, iy, iz) + 0.5 * Fbar[ip];

            t1 = ux9s + uy9s + uz9s;                                   /* :86-101 */
            eqm1 = -11.0 * rho9 + 19.0 * t1;
            eqm2 = omegepsl * rho9 + omegepslj * t1;
            eqm3 = coef1 * ux9;
            eqm4 = coef1 * uy9;
            eqm5 = coef1 * uz9;
            eqm6 = 2.0 * ux9s - uy9s - uz9s;
            eqm7 = omegxx * eqm6;
            eqm8 = uy9s - uz9s;
            eqm9 = omegxx * eqm8;
            eqm10 = ux9 * uy9;
            eqm11 = uy9 * uz9;
            eqm12 = ux9 * uz9;
            eqm13 = 0.0;
            eqm14 = 0.0;
            eqm15 = 0.0;

            sum1 = f9[1] + f9[2] + f9[3] + f9[4] + f9[5] + f9[6];      /* :103-119 */
            sum2 = f9[7] + f9[8] + f9[9] + f9[10] + f9[11] + f9[12]
                 + f9[13] + f9[14] + f9[15] + f9[16] + f9[17] + f9[18];
            sum3 = f9[7] - f9[8] + f9[9] - f9[10] + f9[11] - f9[12]
                 + f9[13] - f9[14];
            sum4 = f9[7] + f9[8] - f9[9] - f9[10] + f9[15] - f9[16]
                 + f9[17] - f9[18];
            sum5 = f9[11] + f9[12] - f9[13] - f9[14] + f9[15] + f9[16]
                 - f9[17] - f9[18];
            sum6 = f9[1] + f9[2];
            sum7 = f9[3] + f9[4] + f9[5] + f9[6];
            sum8 = f9[7] + f9[8] + f9[9] + f9[10] + f9[11] + f9[12]
                 + f9[13] + f9[14];
            sum9 = f9[15] + f9[16] + f9[17] + f9[18];
            sum10 = f9[3] + f9[4] - f9[5] - f9[6];
            sum11 = f9[7] + f9[8] + f9[9] + f9[10] - f9[11] - f9[12]
                  - f9[13] - f9[14];

            evlm1 = -30.0 * f9[0] + coef2 * sum1 + coef3 * sum2;       /* :121-138 */
            evlm2 = 12.0 * f9[0] + coef4 * sum1 + sum2;
            evlm3 = coef4 * (f9[1] - f9[2]) + sum3;
            evlm4 = coef4 * (f9[3] - f9[4]) + sum4;
            evlm5 = coef4 * (f9[5] - f9[6]) + sum5;
            evlm6 = coef5 * sum6 - sum7 + sum8 - coef5 * sum9;
            evlm7 = coef4 * sum6 + coef5 * sum7 + sum8 - coef5 * sum9;
            evlm8 = sum10 + sum11;
            evlm9 = -coef5 * sum10 + sum11;
            evlm10 = f9[7] - f9[8] - f9[9] + f9[10];
            evlm11 = f9[15] - f9[16] - f9[17] + f9[18];
            evlm12 = f9[11] - f9[12] - f9[13] + f9[14];
            evlm13 = f9[7] - f9[8] + f9[9] - f9[10] - f9[11] + f9[12]
                   - f9[13] + f9[14];
            evlm14 = -f9[7] - f9[8] + f9[9] + f9[10] + f9[15] - f9[16]
                   + f9[17] - f9[18];
            evlm15 = f9[11] + f9[12] - f9[13] - f9[14] - f9[15] - f9[16]
                   + f9[17] + f9[18];

            eqmc1 = evlm1 - s1 * (evlm1 - eqm1);                       /* :140-154 */
            eqmc2 = evlm2 - s2 * (evlm2 - eqm2);
            eqmc3 = evlm3 - s4 * (evlm3 - eqm3);
            eqmc4 = evlm4 - s4 * (evlm4 - eqm4);
            eqmc5 = evlm5 - s4 * (evlm5 - eqm5);
            eqmc6 = evlm6 - s9 * (evlm6 - eqm6);
            eqmc7 = evlm7 - s10 * (evlm7 - eqm7);
            eqmc8 = evlm8 - s9 * (evlm8 - eqm8);
            eqmc9 = evlm9 - s10 * (evlm9 - eqm9);
            eqmc10 = evlm10 - s13 * (evlm10 - eqm10);
            eqmc11 = evlm11 - s13 * (evlm11 - eqm11);
            eqmc12 = evlm12 - s13 * (evlm12 - eqm12);
            eqmc13 = evlm13 - s16 * (evlm13 - eqm13);
            eqmc14 = evlm14 - s16 * (evlm14 - eqm14);
            eqmc15 = evlm15 - s16 * (evlm15 - eqm15);

            tl1 = val1i * rho9;                                        /* :157-177 */
            tl2 = coef2 * val2i * eqmc1;
            tl3 = coef3 * val2i * eqmc1;
            tl4 = coef4 * val3i * eqmc2;
            tl5 = val3i * eqmc2;
            tl6 = val4i * ux9;
            tl7 = val5i * eqmc3;
            tl8 = val4i * uy9;
            tl9 = val5i * eqmc4;
            tl10 = val4i * uz9;
            tl11 = val5i * eqmc5;
            tl12 = val6i * eqmc6;
            tl13 = val7i * eqmc7;
            tl14 = val8i * eqmc8;
            tl15 = val9i * eqmc9;
            tl16 = -coef4i * eqmc10;
            tl17 = -coef4i * eqmc11;
            tl18 = -coef4i * eqmc12;
            tl19 = coef3i * eqmc13;
            tl20 = coef3i * eqmc14;
            tl21 = coef3i * eqmc15;

            f9[0] = tl1 - 30.0 * val2i * eqmc1 + val8 * val3i * eqmc2; /* :180 */

            suma = tl1 + tl2 + tl4;                                    /* :182-198 */
            sumb = tl1 + tl3 + tl5;
            sumc = tl6 + coef4 * tl7;
            sumd = coef5 * tl12 + coef4 * tl13;
            sume = tl8 + coef4 * tl9;
            sumf = -tl12 + coef5 * tl13 + tl14 - coef5 * tl15;
            sumg = tl10 + coef4 * tl11;
            sumh = -tl12 + coef5 * tl13 - tl14 + coef5 * tl15;
            sumi = tl12 + tl13 + tl14 + tl15;
            sumk = tl12 + tl13 - tl14 - tl15;
            sump = -coef5 * tl12 - coef5 * tl13;
            sum67 = tl6 + tl7;
            sum89 = tl8 + tl9;
            sum1011 = tl10 + tl11;

            f9[1] = suma + sumc + sumd;                                /* :200-220 */
            f9[2] = suma - sumc + sumd;
            f9[3] = suma + sume + sumf;
            f9[4] = suma - sume + sumf;
            f9[5] = suma + sumg + sumh;
            f9[6] = suma - sumg + sumh;

            f9[7] = sumb + sum67 + sum89 + sumi + tl16 + tl19 - tl20;
            f9[8] = sumb - sum67 + sum89 + sumi - tl16 - tl19 - tl20;
            f9[9] = sumb + sum67 - sum89 + sumi - tl16 + tl19 + tl20;
            f9[10] = sumb - sum67 - sum89 + sumi + tl16 - tl19 + tl20;

            f9[11] = sumb + sum67 + sum1011 + sumk + tl18 - tl19 + tl21;
            f9[12] = sumb - sum67 + sum1011 + sumk - tl18 + tl19 + tl21;
            f9[13] = sumb + sum67 - sum1011 + sumk - tl18 - tl19 - tl21;
            f9[14] = sumb - sum67 - sum1011 + sumk + tl18 + tl19 - tl21;

            f9[15] = sumb + sum89 + sum1011 + sump + tl17 + tl20 - tl21;
            f9[16] = sumb - sum89 + sum1011 + sump - tl17 - tl20 - tl21;
            f9[17] = sumb + sum89 - sum1011 + sump - tl17 + tl20 + tl21;
            f9[18] = sumb - sum89 - sum1011 + sump + tl17 - tl20 + tl21;

            for (ipi = 1; ipi <= 10; ++ipi) {                          /* :224-242 */
                ip = p->ipstay[ipi - 1];
                imove = ix + cix[ip];
                jmove = iy + ciy[ip];
                kmove = iz + ciz[ip];
                if (imove < 1 || imove > lx) {
                    F_(r, ipopp[ip], ix, iy, iz) = f9[ip] + 0.5 * Fbar[ip];
                } else if (jmove < 1) {
                    TY19(r, r->tmpymS, ip, imove, kmove) = f9[ip] + 0.5 * Fbar[ip];
                } else if (jmove > ly) {
                    TY19(r, r->tmpypS, ip, imove, kmove) = f9[ip] + 0.5 * Fbar[ip];
                } else if (kmove < 1) {
                    TZ19(r, r->tmpzmS, ip, imove, jmove) = f9[ip] + 0.5 * Fbar[ip];
                } else if (kmove > lz) {
                    TZ19(r, r->tmpzpS, ip, imove, jmove) = f9[ip] + 0.5 * Fbar[ip];
                } else {
                    F_(r, ipopp[ip], ix, iy, iz) = f9[ip] + 0.5 * Fbar[ip];
                }
            }
        }
        /* 111 continue */
        for (ipi = 1; ipi <= 9; ++ipi) {                               /* :245-264 */
            ip = p->ipswap[ipi - 1];
            imove = ix + cix[ip];
            jmove = iy + ciy[ip];
            kmove = iz + ciz[ip];
            if (imove < 1 || imove > lx) {
                F_(r, ipopp[ip], ix, iy, iz) = f9[ip] + 0.5 * Fbar[ip];
            } else if (jmove < 1) {
                TY19(r, r->tmpymS, ip, imove, kmove) = f9[ip] + 0.5 * Fbar[ip];
            } else if (jmove > ly) {
                TY19(r, r->tmpypS, ip, imove, kmove) = f9[ip] + 0.5 * Fbar[ip];
            } else if (kmove < 1) {
                TZ19(r, r->tmpzmS, ip, imove, jmove) = f9[ip] + 0.5 * Fbar[ip];
            } else if (kmove > lz) {
                TZ19(r, r->tmpzpS, ip, imove, jmove) = f9[ip] + 0.5 * Fbar[ip];
            } else {
                F_(r, ipopp[ip], ix, iy, iz) = F_(r, ip, imove, jmove, kmove);
                F_(r, ip, imove, jmove, kmove) = f9[ip] + 0.5 * Fbar[ip];
            }
        }
    }
}

/* ---- collision.f90:295-305 ----------------------------------------------------------- */
void orc_pack_y(const orc_para *p, orc_rank *r)
{
    const int lx = r->lx, lz = r->lz;
    int ix, k;
    (void)p;
    for (k = 0; k <= lz + 1; ++k)
        for (ix = 1; ix <= lx; ++ix) {
            TY5(r, r->ypS5, 1, ix, k) = TY19(r, r->tmpypS, 3, ix, k);
            TY5(r, r->ypS5, 2, ix, k) = TY19(r, r->tmpypS, 7, ix, k);
            TY5(r, r->ypS5, 3, ix, k) = TY19(r, r->tmpypS, 8, ix, k);
            TY5(r, r->ypS5, 4, ix, k) = TY19(r, r->tmpypS, 15, ix, k);
            TY5(r, r->ypS5, 5, ix, k) = TY19(r, r->tmpypS, 17, ix, k);

            TY5(r, r->ymS5, 1, ix, k) = TY19(r, r->tmpymS, 4, ix, k);
            TY5(r, r->ymS5, 2, ix, k) = TY19(r, r->tmpymS, 9, ix, k);
            TY5(r, r->ymS5, 3, ix, k) = TY19(r, r->tmpymS, 10, ix, k);
            TY5(r, r->ymS5, 4, ix, k) = TY19(r, r->tmpymS, 16, ix, k);
            TY5(r, r->ymS5, 5, ix, k) = TY19(r, r->tmpymS, 18, ix, k);
        }
}

/* ---- collision.f90:309-314: tmpymR(me) = tmpypS(mym), tmpypR(me) = tmpymS(myp) ------- */
void orc_deliver_y(orc_world *w)
{
    int id;
    for (id = 0; id < w->nproc; ++id) {
        orc_rank *r = &w->r[id];
        size_t nb = (size_t)5 * r->lx * (r->lz + 2) * sizeof(double);
        memcpy(r->ymR5, w->r[r->mym].ypS5, nb);   /* recv from mym, tag 0 <- its send to myp, tag 0 */
        memcpy(r->ypR5, w->r[r->myp].ymS5, nb);   /* recv from myp, tag 1 <- its send to mym, tag 1 */
    }
}

/* ---- collision.f90:316-347 ----------------------------------------------------------- */
void orc_unpack_y_pack_z(const orc_para *p, orc_rank *r)
{
    const int lx = r->lx, ly = r->ly, lz = r->lz;
    int ix, iy, iz;
    (void)p;
    for (iz = 1; iz <= lz; ++iz) {
        for (ix = 1; ix <= lx; ++ix) {                                 /* :318-322 */
            F_(r, 3, ix, 1, iz) = TY5(r, r->ymR5, 1, ix, iz);
            if (ix >= 2) F_(r, 7, ix, 1, iz) = TY5(r, r->ymR5, 2, ix, iz);
            if (ix <= lx - 1) F_(r, 8, ix, 1, iz) = TY5(r, r->ymR5, 3, ix, iz);
            F_(r, 15, ix, 1, iz) = TY5(r, r->ymR5, 4, ix, iz);
            F_(r, 17, ix, 1, iz) = TY5(r, r->ymR5, 5, ix, iz);
        }
        for (ix = 1; ix <= lx; ++ix) {                                 /* :324-328 */
            F_(r, 4, ix, ly, iz) = TY5(r, r->ypR5, 1, ix, iz);
            if (ix >= 2) F_(r, 9, ix, ly, iz) = TY5(r, r->ypR5, 2, ix, iz);
            if (ix <= lx - 1) F_(r, 10, ix, ly, iz) = TY5(r, r->ypR5, 3, ix, iz);
            F_(r, 16, ix, ly, iz) = TY5(r, r->ypR5, 4, ix, iz);
            F_(r, 18, ix, ly, iz) = TY5(r, r->ypR5, 5, ix, iz);
        }
    }
    for (ix = 1; ix <= lx; ++ix) {                                     /* :331-334 */
        TZ19(r, r->tmpzmS, 17, ix, 1) = TY5(r, r->ymR5, 5, ix, 0);
        TZ19(r, r->tmpzmS, 18, ix, ly) = TY5(r, r->ypR5, 5, ix, 0);
        TZ19(r, r->tmpzpS, 15, ix, 1) = TY5(r, r->ymR5, 4, ix, lz + 1);
        TZ19(r, r->tmpzpS, 16, ix, ly) = TY5(r, r->ypR5, 4, ix, lz + 1);
    }
    for (iy = 1; iy <= ly; ++iy)                                       /* :337-347 */
        for (ix = 1; ix <= lx; ++ix) {
            TZ5(r, r->zpS5, 1, ix, iy) = TZ19(r, r->tmpzpS, 5, ix, iy);
            TZ5(r, r->zpS5, 2, ix, iy) = TZ19(r, r->tmpzpS, 11, ix, iy);
            TZ5(r, r->zpS5, 3, ix, iy) = TZ19(r, r->tmpzpS, 12, ix, iy);
            TZ5(r, r->zpS5, 4, ix, iy) = TZ19(r, r->tmpzpS, 15, ix, iy);
            TZ5(r, r->zpS5, 5, ix, iy) = TZ19(r, r->tmpzpS, 16, ix, iy);

            TZ5(r, r->zmS5, 1, ix, iy) = TZ19(r, r->tmpzmS, 6, ix, iy);
            TZ5(r, r->zmS5, 2, ix, iy) = TZ19(r, r->tmpzmS, 13, ix, iy);
            TZ5(r, r->zmS5, 3, ix, iy) = TZ19(r, r->tmpzmS, 14, ix, iy);
            TZ5(r, r->zmS5, 4, ix, iy) = TZ19(r, r->tmpzmS, 17, ix, iy);
            TZ5(r, r->zmS5, 5, ix, iy) = TZ19(r, r->tmpzmS, 18, ix, iy);
        }
}

/* ---- collision.f90:351-356: tmpzmR(me) = tmpzpS(mzm), tmpzpR(me) = tmpzmS(mzp) ------- */
void orc_deliver_z(orc_world *w)
{
    int id;
    for (id = 0; id < w->nproc; ++id) {
        orc_rank *r = &w->r[id];
        size_t nb = (size_t)5 * r->lx * r->ly * sizeof(double);
        memcpy(r->zmR5, w->r[r->mzm].zpS5, nb);
        memcpy(r->zpR5, w->r[r->mzp].zmS5, nb);
    }
}

/* ---- collision.f90:358-370 ----------------------------------------------------------- */
void orc_unpack_z(const orc_para *p, orc_rank *r)
{
    const int lx = r->lx, ly = r->ly, lz = r->lz;
    int ix, iy;
    (void)p;
    for (iy = 1; iy <= ly; ++iy)
        for (ix = 1; ix <= lx; ++ix) {
            F_(r, 5, ix, iy, 1) = TZ5(r, r->zmR5, 1, ix, iy);
            if (ix >= 2) F_(r, 11, ix, iy, 1) = TZ5(r, r->zmR5, 2, ix, iy);
            if (ix <= lx - 1) F_(r, 12, ix, iy, 1) = TZ5(r, r->zmR5, 3, ix, iy);
            F_(r, 15, ix, iy, 1) = TZ5(r, r->zmR5, 4, ix, iy);
            F_(r, 16, ix, iy, 1) = TZ5(r, r->zmR5, 5, ix, iy);
        }
    for (iy = 1; iy <= ly; ++iy)
        for (ix = 1; ix <= lx; ++ix) {
            F_(r, 6, ix, iy, lz) = TZ5(r, r->zpR5, 1, ix, iy);
            if (ix >= 2) F_(r, 13, ix, iy, lz) = TZ5(r, r->zpR5, 2, ix, iy);
            if (ix <= lx - 1) F_(r, 14, ix, iy, lz) = TZ5(r, r->zpR5, 3, ix, iy);
            F_(r, 17, ix, iy, lz) = TZ5(r, r->zpR5, 4, ix, iy);
            F_(r, 18, ix, iy, lz) = TZ5(r, r->zpR5, 5, ix, iy);
        }
}

/* ---- collision_MRT on every rank (collision.f90:24-273) ------------------------------ */
static void pf_collide(void *c, int id)
{
    orc_world *w = (orc_world *)c;
    orc_collision_local(&w->p, &w->r[id]);
    orc_pack_y(&w->p, &w->r[id]);
}
static void pf_unpack_y(void *c, int id) { orc_world *w = (orc_world *)c; orc_unpack_y_pack_z(&w->p, &w->r[id]); }
static void pf_unpack_z(void *c, int id) { orc_world *w = (orc_world *)c; orc_unpack_z(&w->p, &w->r[id]); }

void orc_collision_MRT(orc_world *w)
{
    orc_parallel_for(w->nproc, pf_collide, w);
    orc_deliver_y(w);
    orc_parallel_for(w->nproc, pf_unpack_y, w);
    orc_deliver_z(w);
    orc_parallel_for(w->nproc, pf_unpack_z, w);
}

/* ---- collision.f90:378-463 ----------------------------------------------------------- */
static void pf_macrovar(void *c, int id);
void orc_macrovar(orc_world *w) { orc_parallel_for(w->nproc, pf_macrovar, w); }

static void pf_macrovar(void *c, int id)
{
    orc_world *w = (orc_world *)c;
    const orc_para *p = &w->p;
    const int nyh = p->ny / 2, nzh = p->nz / 2;   /* var_inc.f90:54 */
    {
        orc_rank *r = &w->r[id];
        int ix, iy, iz, ip;
        for (iz = 1; iz <= r->lz; ++iz)
        for (iy = 1; iy <= r->ly; ++iy)
        for (ix = 1; ix <= r->lx; ++ix) {
            if (IB_(r, ix, iy, iz) < 0) {                              /* :394 */
                double f9[NPOP];
                double sum1, sum2, sum3, sum4, sum5, sum6, ux9, uy9, uz9, rho9;
                for (ip = 0; ip < NPOP; ++ip) f9[ip] = F_(r, ip, ix, iy, iz);
                sum1 = f9[7] - f9[10];                                 /* :398-405 */
                sum2 = f9[9] - f9[8];
                sum3 = f9[11] - f9[14];
                sum4 = f9[13] - f9[12];
                sum5 = f9[15] - f9[18];
                sum6 = f9[17] - f9[16];
                ux9 = f9[1] - f9[2] + sum1 + sum2 + sum3 + sum4;       /* :407-409 */
                uy9 = f9[3] - f9[4] + sum1 - sum2 + sum5 + sum6;
                uz9 = f9[5] - f9[6] + sum3 - sum4 + sum5 - sum6;
                rho9 = f9[0] + f9[1] + f9[2] + f9[3] + f9[4] + f9[5] + f9[6]   /* :411-413 */
                     + f9[7] + f9[8] + f9[9] + f9[10] + f9[11] + f9[12]
                     + f9[13] + f9[14] + f9[15] + f9[16] + f9[17] + f9[18];
                S_(r, r->ux, ix, iy, iz) = ux9 + S_(r, r->force_realx, ix, iy, iz) / 2.;  /* :415-418 */
                S_(r, r->uy, ix, iy, iz) = uy9 + S_(r, r->force_realy, ix, iy, iz) / 2.;
                S_(r, r->uz, ix, iy, iz) = uz9 + S_(r, r->force_realz, ix, iy, iz) / 2.;
                S_(r, r->rho, ix, iy, iz) = rho9;
            } else if (p->ipart) {                                     /* :420-459 */
                int id1 = S_(r, r->isnodes, ix, iy, iz);               /* 1-based particle id */
                double xpnt = (double)ix - 0.5;
                double ypnt = (double)iy - 0.5 + r->globaly;
                double zpnt = (double)iz - 0.5 + r->globalz;
                double xc = w->ypglb[0 + 3 * (id1 - 1)];
                double yc = w->ypglb[1 + 3 * (id1 - 1)];
                double zc = w->ypglb[2 + 3 * (id1 - 1)];
                double xx0, yy0, zz0, w1, w2, w3, omg1, omg2, omg3;
                if ((yc - ypnt) > (double)nyh) yc = yc - (double)p->ny;
                if ((yc - ypnt) < -(double)nyh) yc = yc + (double)p->ny;
                if ((zc - zpnt) > (double)nzh) zc = zc - (double)p->nz;
                if ((zc - zpnt) < -(double)nzh) zc = zc + (double)p->nz;
                xx0 = xpnt - xc; yy0 = ypnt - yc; zz0 = zpnt - zc;
                w1 = w->wp[0 + 3 * (id1 - 1)]; w2 = w->wp[1 + 3 * (id1 - 1)]; w3 = w->wp[2 + 3 * (id1 - 1)];
                omg1 = w->omgp[0 + 3 * (id1 - 1)]; omg2 = w->omgp[1 + 3 * (id1 - 1)]; omg3 = w->omgp[2 + 3 * (id1 - 1)];
                S_(r, r->ux, ix, iy, iz) = w1 + (omg2 * zz0 - omg3 * yy0);
                S_(r, r->uy, ix, iy, iz) = w2 + (omg3 * xx0 - omg1 * zz0);
                S_(r, r->uz, ix, iy, iz) = w3 + (omg1 * yy0 - omg2 * xx0);
                S_(r, r->rho, ix, iy, iz) = p->rhopart;
            }
        }
    }
}

/* ---- saveload.f90:3929-4054 -- vortcalc + exchng8 ------------------------------------------
 * Vorticity of the velocity field held in ux,uy,uz (the caller ran macrovar): central
 * differences, one-sided at the channel walls where the no-slip wall sits half a spacing
 * outside the first node (:3971-3980), neighbours across a subdomain face from the planes
 * exchng8 (:4027-4054) brings in: tmp?B = the +z neighbour's plane 1, tmp?F = the -z neighbour's
 * plane lz, tmp?R = the +y neighbour's row 1, tmp?L = the -y neighbour's row ly.  A node inside
 * a particle gets twice the particle's angular velocity (:4011-4020).  Results go to global
 * arrays in the (lx,ny,nz) layout. */
void orc_vortcalc(const orc_world *w, double *oxg, double *oyg, double *ozg)
{
    const orc_para *p = &w->p;
    int id;
    for (id = 0; id < w->nproc; ++id) {
        const orc_rank *r = &w->r[id];
        const orc_rank *zp = &w->r[r->mzp], *zm = &w->r[r->mzm], *yp = &w->r[r->myp], *ym = &w->r[r->mym];
        const int lx = r->lx, ly = r->ly, lz = r->lz;
        int i, j, k;
#define U_(rr, a, i, j, k) ((rr)->a[(size_t)((i)-1) + (size_t)(rr)->lx * ((size_t)((j)-1) + (size_t)(rr)->ly * (size_t)((k)-1))])
        for (k = 1; k <= lz; ++k)
        for (j = 1; j <= ly; ++j)
        for (i = 1; i <= lx; ++i) {
            double ox, oy, oz;
            if (IB_(r, i, j, k) < 0) {                                         /* :3969 */
                double pwx, pvx, pwy, puy, puz, pvz;
                if (i == 1) {                                                  /* :3971-3973 */
                    pwx = (3.0 * U_(r, uz, i, j, k) + U_(r, uz, i + 1, j, k)) / 3.0;
                    pvx = (3.0 * U_(r, uy, i, j, k) + U_(r, uy, i + 1, j, k)) / 3.0;
                } else if (i == lx) {                                          /* :3974-3976 */
                    pwx = -(3.0 * U_(r, uz, i, j, k) + U_(r, uz, i - 1, j, k)) / 3.0;
                    pvx = -(3.0 * U_(r, uy, i, j, k) + U_(r, uy, i - 1, j, k)) / 3.0;
                } else {                                                       /* :3977-3980 */
                    pwx = (U_(r, uz, i + 1, j, k) - U_(r, uz, i - 1, j, k)) / 2.0;
                    pvx = (U_(r, uy, i + 1, j, k) - U_(r, uy, i - 1, j, k)) / 2.0;
                }
                if (j == 1) {                                                  /* :3982-3984: tmpu?L = mym's row ly */
                    pwy = (U_(r, uz, i, j + 1, k) - U_(ym, uz, i, ym->ly, k)) / 2.0;
                    puy = (U_(r, ux, i, j + 1, k) - U_(ym, ux, i, ym->ly, k)) / 2.0;
                } else if (j == ly) {                                          /* :3985-3987: tmpu?R = myp's row 1 */
                    pwy = (U_(yp, uz, i, 1, k) - U_(r, uz, i, j - 1, k)) / 2.0;
                    puy = (U_(yp, ux, i, 1, k) - U_(r, ux, i, j - 1, k)) / 2.0;
                } else {
                    pwy = (U_(r, uz, i, j + 1, k) - U_(r, uz, i, j - 1, k)) / 2.0;
                    puy = (U_(r, ux, i, j + 1, k) - U_(r, ux, i, j - 1, k)) / 2.0;
                }
                if (k == 1) {                                                  /* :3993-3995: tmpu?F = mzm's plane lz */
                    puz = (U_(r, ux, i, j, k + 1) - U_(zm, ux, i, j, zm->lz)) / 2.0;
                    pvz = (U_(r, uy, i, j, k + 1) - U_(zm, uy, i, j, zm->lz)) / 2.0;
                } else if (k == lz) {                                          /* :3996-3998: tmpu?B = mzp's plane 1 */
                    puz = (U_(zp, ux, i, j, 1) - U_(r, ux, i, j, k - 1)) / 2.0;
                    pvz = (U_(zp, uy, i, j, 1) - U_(r, uy, i, j, k - 1)) / 2.0;
                } else {
                    puz = (U_(r, ux, i, j, k + 1) - U_(r, ux, i, j, k - 1)) / 2.0;
                    pvz = (U_(r, uy, i, j, k + 1) - U_(r, uy, i, j, k - 1)) / 2.0;
                }
                ox = pwy - pvz;                                                /* :4004-4006 */
                oy = puz - pwx;
                oz = pvx - puy;
            } else {                                                           /* :4008-4019 */
                const int id1 = S_(r, r->isnodes, i, j, k);
                ox = 2.0 * w->omgp[0 + 3 * (id1 - 1)];
                oy = 2.0 * w->omgp[1 + 3 * (id1 - 1)];
                oz = 2.0 * w->omgp[2 + 3 * (id1 - 1)];
            }
            {
                const size_t g = (size_t)(i - 1) + (size_t)p->nx * ((size_t)(j - 1 + r->globaly) + (size_t)p->ny * (size_t)(k - 1 + r->globalz));
                oxg[g] = ox; oyg[g] = oy; ozg[g] = oz;
            }
        }
#undef U_
    }
}

/* ---- saveload.f90:2031-2091 -- local strain rate from the non-equilibrium moments (sijstat00) ------
 * Sij*Sij of every fluid node (ibnodes < 0) from f (post-streaming) and the rho,u arrays as macrovar
 * left them (Yu et al., Computers & Fluids 35 (2006) 957, appendix): the six second-order moments
 * minus their equilibria, times their relaxation rates, are the strain-rate tensor up to constants --
 * no finite differences, no halo.  The sums are collision_MRT's (:104-138).  Global (nx,ny,nz) layout;
 * solid nodes are left untouched (the reference's automatic array is undefined there). */
void orc_sijstat(const orc_world *w, double *sij2g)
{
    const orc_para *p = &w->p;
    const double coef2 = p->coef2, coef3 = p->coef3, coef5 = p->coef5, s1 = p->s1, s9 = p->s9;
    int id;
    for (id = 0; id < w->nproc; ++id) {
        const orc_rank *r = &w->r[id];
        int ix, iy, iz;
        for (iz = 1; iz <= r->lz; ++iz)
        for (iy = 1; iy <= r->ly; ++iy)
        for (ix = 1; ix <= r->lx; ++ix) {
            const double *f9;
            double rho9, ux9, uy9, uz9, ux9s, uy9s, uz9s, eqm1, eqm6, eqm8, eqm10, eqm11, eqm12;
            double sum1, sum2, sum6, sum7, sum8, sum9, sum10, sum11, evlm1, evlm6, evlm8, evlm10, evlm11, evlm12;
            double neqm1, neqm9, neqm11, neqm13, neqm14, neqm15, Sxx, Syy, Szz, Sxy, Syz, Szx;
            if (!(IB_(r, ix, iy, iz) < 0)) continue;                                  /* :2034 */
            f9 = &F_(r, 0, ix, iy, iz);                                               /* :2036 */
            rho9 = S_(r, r->rho, ix, iy, iz);
            ux9 = S_(r, r->ux, ix, iy, iz); uy9 = S_(r, r->uy, ix, iy, iz); uz9 = S_(r, r->uz, ix, iy, iz);
            ux9s = ux9 * ux9; uy9s = uy9 * uy9; uz9s = uz9 * uz9;
            eqm1 = -(11.0 * rho9) + 19.0 * ((ux9s + uy9s) + uz9s);                    /* :2046-2051 */
            eqm6 = (2.0 * ux9s - uy9s) - uz9s;
            eqm8 = uy9s - uz9s;
            eqm10 = ux9 * uy9; eqm11 = uy9 * uz9; eqm12 = ux9 * uz9;
            sum1 = ((((f9[1] + f9[2]) + f9[3]) + f9[4]) + f9[5]) + f9[6];             /* :2053-2064 */
            sum2 = ((((((((((f9[7] + f9[8]) + f9[9]) + f9[10]) + f9[11]) + f9[12]) + f9[13]) + f9[14]) + f9[15]) + f9[16])
                    + f9[17]) + f9[18];
            sum6 = f9[1] + f9[2];
            sum7 = ((f9[3] + f9[4]) + f9[5]) + f9[6];
            sum8 = ((((((f9[7] + f9[8]) + f9[9]) + f9[10]) + f9[11]) + f9[12]) + f9[13]) + f9[14];
            sum9 = ((f9[15] + f9[16]) + f9[17]) + f9[18];
            sum10 = ((f9[3] + f9[4]) - f9[5]) - f9[6];
            sum11 = ((((((f9[7] + f9[8]) + f9[9]) + f9[10]) - f9[11]) - f9[12]) - f9[13]) - f9[14];
            evlm1 = (-(30.0 * f9[0]) + coef2 * sum1) + coef3 * sum2;                  /* :2066-2071 */
            evlm6 = ((coef5 * sum6 - sum7) + sum8) - coef5 * sum9;
            evlm8 = sum10 + sum11;
            evlm10 = ((f9[7] - f9[8]) - f9[9]) + f9[10];
            evlm11 = ((f9[15] - f9[16]) - f9[17]) + f9[18];
            evlm12 = ((f9[11] - f9[12]) - f9[13]) + f9[14];
            neqm1 = s1 * (evlm1 - eqm1);                                              /* :2073-2078 */
            neqm9 = s9 * (evlm6 - eqm6);
            neqm11 = s9 * (evlm8 - eqm8);
            neqm13 = s9 * (evlm10 - eqm10);
            neqm14 = s9 * (evlm11 - eqm11);
            neqm15 = s9 * (evlm12 - eqm12);
            Sxx = -((neqm1 + 19.0 * neqm9) / 38.0);                                   /* :2080-2085 */
            Syy = -((2.0 * neqm1 - 19.0 * (neqm9 - 3.0 * neqm11)) / 76.0);
            Szz = -((2.0 * neqm1 - 19.0 * (neqm9 + 3.0 * neqm11)) / 76.0);
            Sxy = -(1.5 * neqm13); Syz = -(1.5 * neqm14); Szx = -(1.5 * neqm15);
            sij2g[(size_t)(ix - 1) + (size_t)p->nx * ((size_t)(iy - 1 + r->globaly) + (size_t)p->ny * (size_t)(iz - 1 + r->globalz))] =
                ((Sxx * Sxx + Syy * Syy) + Szz * Szz) + 2.0 * ((Sxy * Sxy + Syz * Syz) + Szx * Szx);   /* :2087-2088 */
        }
    }
}

/* ---- collision.f90:469-480 ----------------------------------------------------------- */
static void pf_rhoupdat(void *c, int id);
void orc_rhoupdat(orc_world *w) { orc_parallel_for(w->nproc, pf_rhoupdat, w); }

static void pf_rhoupdat(void *c, int id)
{
    orc_world *w = (orc_world *)c;
    {
        orc_rank *r = &w->r[id];
        const size_t n3 = (size_t)r->lx * r->ly * r->lz;
        size_t n;
        int ip;
        for (n = 0; n < n3; ++n) {
            double rho = r->f[0 + NPOP * n];
            for (ip = 1; ip <= NPOP - 1; ++ip) rho = rho + r->f[ip + NPOP * n];
            r->rho[n] = rho;
        }
    }
}

/* ---- collision.f90:487-513 ----------------------------------------------------------- */
double orc_avedensity(orc_world *w, int64_t *nfluidtotal_out)
{
    int id;
    int64_t nfluidtotal = 0;
    double rhomean = 0.0;
    for (id = 0; id < w->nproc; ++id) {          /* per-rank partials, then "Allreduce" in rank order */
        orc_rank *r = &w->r[id];
        int ix, iy, iz;
        int64_t nfluid0 = 0;
        double rhomean0 = 0.0;
        for (iz = 1; iz <= r->lz; ++iz)
        for (iy = 1; iy <= r->ly; ++iy)
        for (ix = 1; ix <= r->lx; ++ix)
            if (IB_(r, ix, iy, iz) < 0) {
                nfluid0 += 1;                                          /* :497 */
                rhomean0 += S_(r, r->rho, ix, iy, iz);                 /* :498 */
            }
        nfluidtotal += nfluid0;                                        /* :500 */
        rhomean += rhomean0;                                           /* :501 */
    }
    rhomean = rhomean / (double)nfluidtotal;                           /* :503 */
    for (id = 0; id < w->nproc; ++id) {                                /* :505-511 */
        orc_rank *r = &w->r[id];
        const size_t n3 = (size_t)r->lx * r->ly * r->lz;
        size_t n;
        for (n = 0; n < n3; ++n) r->rho[n] = r->rho[n] - rhomean;
    }
    if (nfluidtotal_out) *nfluidtotal_out = nfluidtotal;
    return rhomean;
}

/* ---- test conveniences (not in the reference) ---------------------------------------- */
void orc_gather_f(const orc_world *w, double *fg)
{
    const int nx = w->p.nx, ny = w->p.ny;
    int id;
    for (id = 0; id < w->nproc; ++id) {
        const orc_rank *r = &w->r[id];
        int ix, iy, iz;
        for (iz = 1; iz <= r->lz; ++iz)
        for (iy = 1; iy <= r->ly; ++iy)
        for (ix = 1; ix <= r->lx; ++ix) {
            size_t g = (size_t)(ix - 1) + (size_t)nx * ((size_t)(iy - 1 + r->globaly) + (size_t)ny * (size_t)(iz - 1 + r->globalz));
            memcpy(&fg[NPOP * g], &F_(r, 0, ix, iy, iz), NPOP * sizeof(double));
        }
    }
}

void orc_scatter_f(orc_world *w, const double *fg)
{
    const int nx = w->p.nx, ny = w->p.ny;
    int id;
    for (id = 0; id < w->nproc; ++id) {
        orc_rank *r = &w->r[id];
        int ix, iy, iz;
        for (iz = 1; iz <= r->lz; ++iz)
        for (iy = 1; iy <= r->ly; ++iy)
        for (ix = 1; ix <= r->lx; ++ix) {
            size_t g = (size_t)(ix - 1) + (size_t)nx * ((size_t)(iy - 1 + r->globaly) + (size_t)ny * (size_t)(iz - 1 + r->globalz));
            memcpy(&F_(r, 0, ix, iy, iz), &fg[NPOP * g], NPOP * sizeof(double));
        }
    }
}

static double *orc_pick(const orc_rank *r, int which)
{
    switch (which) {
    case 0: return r->rho;
    case 1: return r->ux;
    case 2: return r->uy;
    case 3: return r->uz;
    case 4: return r->force_realx;
    case 5: return r->force_realy;
    default: return r->force_realz;
    }
}

void orc_gather_scalar(const orc_world *w, int which, double *ag)
{
    const int nx = w->p.nx, ny = w->p.ny;
    int id;
    for (id = 0; id < w->nproc; ++id) {
        const orc_rank *r = &w->r[id];
        const double *a = orc_pick(r, which);
        int iy, iz;
        for (iz = 1; iz <= r->lz; ++iz)
        for (iy = 1; iy <= r->ly; ++iy) {
            size_t g = (size_t)nx * ((size_t)(iy - 1 + r->globaly) + (size_t)ny * (size_t)(iz - 1 + r->globalz));
            memcpy(&ag[g], &S_(r, a, 1, iy, iz), (size_t)r->lx * sizeof(double));
        }
    }
}

void orc_scatter_scalar(orc_world *w, int which, const double *ag)
{
    const int nx = w->p.nx, ny = w->p.ny;
    int id;
    for (id = 0; id < w->nproc; ++id) {
        orc_rank *r = &w->r[id];
        double *a = orc_pick(r, which);
        int iy, iz;
        for (iz = 1; iz <= r->lz; ++iz)
        for (iy = 1; iy <= r->ly; ++iy) {
            size_t g = (size_t)nx * ((size_t)(iy - 1 + r->globaly) + (size_t)ny * (size_t)(iz - 1 + r->globalz));
            memcpy(&S_(r, a, 1, iy, iz), &ag[g], (size_t)r->lx * sizeof(double));
        }
    }
}

/* ib_global: (nx,ny,nz) un-ghosted mask; ghosts are filled periodically in y,z and left
 * fluid (-1) beyond the x walls, which is all the path reads (collision.f90:54,394). */
void orc_scatter_ibnodes(orc_world *w, const int32_t *ibg, const int32_t *isg)
{
    const int nx = w->p.nx, ny = w->p.ny, nz = w->p.nz;
    int id;
    for (id = 0; id < w->nproc; ++id) {
        orc_rank *r = &w->r[id];
        int ix, iy, iz;
        for (iz = 0; iz <= r->lz + 1; ++iz)
        for (iy = 0; iy <= r->ly + 1; ++iy)
        for (ix = 1; ix <= r->lx; ++ix) {
            int gy = ((iy - 1 + r->globaly) % ny + ny) % ny;
            int gz = ((iz - 1 + r->globalz) % nz + nz) % nz;
            size_t g = (size_t)(ix - 1) + (size_t)nx * ((size_t)gy + (size_t)ny * (size_t)gz);
            IB_(r, ix, iy, iz) = ibg[g];
            if (isg && iy >= 1 && iy <= r->ly && iz >= 1 && iz <= r->lz)
                S_(r, r->isnodes, ix, iy, iz) = isg[g];
        }
    }
}

/* ---- accessors for the ctypes wrapper (oracle/oracle.py) ----------------------------- */
void orc_rank_dims(const orc_world *w, int id, int *out9)
{
    const orc_rank *r = &w->r[id];
    out9[0] = r->lx; out9[1] = r->ly; out9[2] = r->lz;
    out9[3] = r->globaly; out9[4] = r->globalz;
    out9[5] = r->mym; out9[6] = r->myp; out9[7] = r->mzm; out9[8] = r->mzp;
}

/* which: 0 f, 1 rho, 2 ux, 3 uy, 4 uz, 5..7 force, 8..11 y 5-slot S-/S+/R-/R+, 12..15 z ditto */
double *orc_rank_array(orc_world *w, int id, int which)
{
    orc_rank *r = &w->r[id];
    switch (which) {
    case 0: return r->f;
    case 1: return r->rho;
    case 2: return r->ux;
    case 3: return r->uy;
    case 4: return r->uz;
    case 5: return r->force_realx;
    case 6: return r->force_realy;
    case 7: return r->force_realz;
    case 8: return r->ymS5;
    case 9: return r->ypS5;
    case 10: return r->ymR5;
    case 11: return r->ypR5;
    case 12: return r->zmS5;
    case 13: return r->zpS5;
    case 14: return r->zmR5;
    case 15: return r->zpR5;
    default: return 0;
    }
}

int32_t *orc_rank_ibnodes(orc_world *w, int id) { return w->r[id].ibnodes; }
const orc_para *orc_world_para(const orc_world *w) { return &w->p; }
int orc_world_nproc(const orc_world *w) { return w->nproc; }
void orc_world_set_para(orc_world *w, const orc_para *p) { w->p = *p; }

void orc_world_set_particles(orc_world *w, int npart, const double *ypglb, const double *wp, const double *omgp)
{
    free(w->ypglb); free(w->wp); free(w->omgp);
    w->npart = npart;
    w->ypglb = (double *)xcalloc((size_t)3 * npart, sizeof(double));
    w->wp = (double *)xcalloc((size_t)3 * npart, sizeof(double));
    w->omgp = (double *)xcalloc((size_t)3 * npart, sizeof(double));
    memcpy(w->ypglb, ypglb, (size_t)3 * npart * sizeof(double));
    memcpy(w->wp, wp, (size_t)3 * npart * sizeof(double));
    memcpy(w->omgp, omgp, (size_t)3 * npart * sizeof(double));
}

/* One rank's local phases, for the gloo test that moves the buffers itself. */
void orc_rank_collision_local(orc_world *w, int id) { orc_collision_local(&w->p, &w->r[id]); orc_pack_y(&w->p, &w->r[id]); }
void orc_rank_unpack_y_pack_z(orc_world *w, int id) { orc_unpack_y_pack_z(&w->p, &w->r[id]); }
void orc_rank_unpack_z(orc_world *w, int id) { orc_unpack_z(&w->p, &w->r[id]); }
