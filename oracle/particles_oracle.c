/*
 * particles_oracle.c -- CPU statement of the particle path (solid mask, boundary links,
 * interpolated bounce-back, momentum-exchange force, refill, lubrication, rigid-body update).
 *
 * TEST INFRASTRUCTURE ONLY (see d3q19_oracle.h).
 *
 * PARITY UNPINNED: the reference snapshot does NOT contain its particle library -- main.f90:64
 * names partlib.f90, no such file exists and no beads_* subroutine is defined anywhere
 * (SURVEY.md fact 2, Appendix B).  What follows is therefore not a restatement of reference code
 * but of the published algorithms the reference's data structures point to:
 *   - solid mask / links: node (ix,iy,iz) sits at (ix-0.5, iy-0.5, iz-0.5) (collision.f90:424-426),
 *     is solid iff |r - r_c| < rad (var_inc.f90:67), periodic images in y,z only
 *     (collision.f90:433-440); ibnodes -1 fluid / >0 solid (para.f90:447, saveload.f90:1359),
 *     isnodes = owning particle id (para.f90:485); links as (xlink,ylink,zlink,iplink,mlink,alink)
 *     (var_inc.f90:127-128);
 *   - interpolated bounce-back: Bouzidi, Firdaouss & Lallemand (2001) linear scheme with the
 *     moving-wall term of Lallemand & Luo (2003), falling back to half-way bounce-back where the
 *     second fluid node is missing;
 *   - hydrodynamic force: Galilean-invariant momentum exchange (Wen et al. 2014; Peng et al. 2016);
 *   - refill of uncovered nodes: equilibrium at the particle's surface velocity and the averaged
 *     neighbour density plus the non-equilibrium part of the fluid neighbour nearest the outward
 *     normal (Caiazzo 2008; Peng et al. 2016);
 *   - short-range repulsion: Feng & Michaelides (2005) eq. (28) form with the stiffness
 *     constants the reference keeps (stf0, stf1, para.f90:356-360; mingap var_inc.f90:67);
 *   - rigid-body update: Newton-Euler with the force averaged over two steps
 *     (forcep/forcepp, torqp/torqpp, para.f90:464-469).
 * It is the checker of the CUDA particle kernels: integer artefacts (mask, link list) must agree
 * bit for bit, populations and forces to rounding.
 *
 * Arrays: populations are the canonical post-streaming f(0:18,nx,ny,nz) of one whole (undecomposed)
 * channel, f[ip + 19*((ix-1) + nx*((iy-1) + ny*(iz-1)))]; own[(ix-1) + nx*((iy-1) + ny*(iz-1))]
 * is the owning particle id (1-based) or -1.  After collision_MRT with solid nodes skipped the
 * post-collision population f*_i(x_f) that streamed into a solid neighbour x_s = x_f + c_i is
 * parked in f(i, x_s) (collision.f90:54,244-264: solid nodes keep running the swap loop), which is
 * where the interpolation reads it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NPOP 19
static const int CX[NPOP] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
static const int CY[NPOP] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
static const int CZ[NPOP] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
static const int OPP[NPOP] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};

typedef struct po_geom {
    int nx, ny, nz;
    double rad;
} po_geom;

static inline int wrap(int j, int n) { int r = (j - 1) % n; if (r < 0) r += n; return r + 1; }
static inline size_t nidx(const po_geom *g, int ix, int iy, int iz)
{
    return (size_t)(ix - 1) + (size_t)g->nx * ((size_t)(iy - 1) + (size_t)g->ny * (size_t)(iz - 1));
}

/* bounding box of particle p in UNWRAPPED node indices: every solid node and every fluid node
 * with a link into p lies inside */
static void bbox(const po_geom *g, const double *c, int lo[3], int hi[3])
{
    for (int d = 0; d < 3; ++d) {
        lo[d] = (int)floor(c[d] - g->rad + 0.5) - 1;
        hi[d] = (int)ceil(c[d] + g->rad + 0.5) + 1;
    }
    if (lo[0] < 1) lo[0] = 1;
    if (hi[0] > g->nx) hi[0] = g->nx;
}

static inline double dist2(const double *c, int jx, int jy, int jz)
{
    const double dx = ((double)jx - 0.5) - c[0], dy = ((double)jy - 0.5) - c[1], dz = ((double)jz - 0.5) - c[2];
    return dx * dx + dy * dy + dz * dz;
}

/* ---- beads_links, part 1: the solid mask ----------------------------------------------------- */
void po_build_mask(const po_geom *g, int npart, const double *ypglb, int32_t *own)
{
    const size_t n = (size_t)g->nx * g->ny * g->nz;
    const double r2 = g->rad * g->rad;
    for (size_t i = 0; i < n; ++i) own[i] = -1;
    for (int p = npart - 1; p >= 0; --p) {             /* the lowest id wins where particles overlap */
        const double *c = ypglb + 3 * p;
        int lo[3], hi[3];
        bbox(g, c, lo, hi);
        for (int jz = lo[2]; jz <= hi[2]; ++jz)
            for (int jy = lo[1]; jy <= hi[1]; ++jy)
                for (int jx = lo[0]; jx <= hi[0]; ++jx)
                    if (dist2(c, jx, jy, jz) < r2) own[nidx(g, jx, wrap(jy, g->ny), wrap(jz, g->nz))] = p + 1;
    }
}

/* ---- beads_links, part 2: boundary links -------------------------------------------------------
 * order: particle, then its box in z,y,x order, then direction 1..18.  q in (0,1] is the fraction
 * of the link between the fluid node and the wall.  Returns the number of links (may exceed
 * maxlink: then only the first maxlink are stored). */
long po_build_links(const po_geom *g, int npart, const double *ypglb, const int32_t *own, long maxlink,
                    int32_t *lx, int32_t *ly, int32_t *lz, int32_t *lip, int32_t *lpart, double *lq)
{
    long n = 0;
    const double r2 = g->rad * g->rad;
    for (int p = 0; p < npart; ++p) {
        const double *c = ypglb + 3 * p;
        int lo[3], hi[3];
        bbox(g, c, lo, hi);
        for (int jz = lo[2]; jz <= hi[2]; ++jz)
            for (int jy = lo[1]; jy <= hi[1]; ++jy)
                for (int jx = lo[0]; jx <= hi[0]; ++jx) {
                    const int iy = wrap(jy, g->ny), iz = wrap(jz, g->nz);
                    if (own[nidx(g, jx, iy, iz)] > 0) continue;                 /* fluid nodes only */
                    for (int ip = 1; ip < NPOP; ++ip) {
                        const int kx = jx + CX[ip], ky = jy + CY[ip], kz = jz + CZ[ip];
                        if (kx < 1 || kx > g->nx) continue;                    /* channel wall, not a particle */
                        if (own[nidx(g, kx, wrap(ky, g->ny), wrap(kz, g->nz))] != p + 1) continue;
                        /* |x_f + t c - r_c|^2 = rad^2, smallest root */
                        const double dx = ((double)jx - 0.5) - c[0], dy = ((double)jy - 0.5) - c[1], dz = ((double)jz - 0.5) - c[2];
                        const double a = (double)(CX[ip] * CX[ip] + CY[ip] * CY[ip] + CZ[ip] * CZ[ip]);
                        const double b = 2.0 * ((double)CX[ip] * dx + (double)CY[ip] * dy + (double)CZ[ip] * dz);
                        const double cc = dx * dx + dy * dy + dz * dz - r2;
                        double disc = b * b - 4.0 * a * cc;
                        if (disc < 0.0) disc = 0.0;
                        double q = (-b - sqrt(disc)) / (2.0 * a);
                        if (q < 0.0) q = 0.0;
                        if (q > 1.0) q = 1.0;
                        if (n < maxlink) {
                            lx[n] = jx; ly[n] = iy; lz[n] = iz; lip[n] = ip; lpart[n] = p + 1; lq[n] = q;
                        }
                        ++n;
                    }
                }
    }
    return n;
}

/* ---- beads_collision: interpolated bounce-back + momentum exchange -------------------------------
 * f: canonical populations after collision_MRT (solid nodes skipped).  fHIp, torqp (3,npart) are
 * overwritten with this step's hydrodynamic force and torque. */
void po_ibb(const po_geom *g, double *f, const int32_t *own, long nlink, const int32_t *lx, const int32_t *ly,
            const int32_t *lz, const int32_t *lip, const int32_t *lpart, const double *lq, int npart,
            const double *ypglb, const double *wp, const double *omgp, double rho0, double *fHIp, double *torqp)
{
    memset(fHIp, 0, sizeof(double) * 3 * (size_t)npart);
    memset(torqp, 0, sizeof(double) * 3 * (size_t)npart);
    for (long l = 0; l < nlink; ++l) {
        const int ix = lx[l], iy = ly[l], iz = lz[l], ip = lip[l], io = OPP[ip], p = lpart[l] - 1;
        const double q = lq[l];
        const double ww = ip <= 6 ? 1.0 / 18.0 : 1.0 / 36.0;
        const size_t nf = nidx(g, ix, iy, iz);
        const int sx = ix + CX[ip], sy = wrap(iy + CY[ip], g->ny), sz = wrap(iz + CZ[ip], g->nz);
        const int bx = ix - CX[ip], by = wrap(iy - CY[ip], g->ny), bz = wrap(iz - CZ[ip], g->nz);
        const size_t ns = nidx(g, sx, sy, sz);
        /* wall point relative to the particle centre, nearest image */
        double c[3] = {ypglb[3 * p], ypglb[3 * p + 1], ypglb[3 * p + 2]};
        double xf[3] = {(double)ix - 0.5, (double)iy - 0.5, (double)iz - 0.5};
        if (c[1] - xf[1] > 0.5 * g->ny) c[1] -= g->ny;
        if (c[1] - xf[1] < -0.5 * g->ny) c[1] += g->ny;
        if (c[2] - xf[2] > 0.5 * g->nz) c[2] -= g->nz;
        if (c[2] - xf[2] < -0.5 * g->nz) c[2] += g->nz;
        const double rx = xf[0] + q * CX[ip] - c[0], ry = xf[1] + q * CY[ip] - c[1], rz = xf[2] + q * CZ[ip] - c[2];
        const double uwx = wp[3 * p] + (omgp[3 * p + 1] * rz - omgp[3 * p + 2] * ry);
        const double uwy = wp[3 * p + 1] + (omgp[3 * p + 2] * rx - omgp[3 * p] * rz);
        const double uwz = wp[3 * p + 2] + (omgp[3 * p] * ry - omgp[3 * p + 1] * rx);
        /* moving-wall term for the population coming back along opp(ip) */
        const double delta = 6.0 * ww * rho0 * (-(CX[ip] * uwx + CY[ip] * uwy + CZ[ip] * uwz));
        const double fs_i = f[ip + NPOP * ns];                     /* f*_i(x_f), parked in the solid node */
        double fnew;
        if (q >= 0.5) {
            const size_t nb = (bx >= 1 && bx <= g->nx) ? nidx(g, bx, by, bz) : nf;
            /* f*_opp(x_f) streamed to x_f - c_i; at a channel wall it bounced back into f(ip, x_f) */
            const double fs_o = (bx >= 1 && bx <= g->nx) ? f[io + NPOP * nb] : f[ip + NPOP * nf];
            const double i2q = 1.0 / (2.0 * q);
            fnew = i2q * fs_i + (2.0 * q - 1.0) * i2q * fs_o + i2q * delta;
        } else {
            const int have_ff = bx >= 1 && bx <= g->nx && own[nidx(g, bx, by, bz)] < 0;
            if (have_ff) fnew = 2.0 * q * fs_i + (1.0 - 2.0 * q) * f[ip + NPOP * nf] + delta;   /* f*_i(x_f - c_i) = f(ip, x_f) */
            else fnew = fs_i + delta;
        }
        f[io + NPOP * nf] = fnew;
        /* Galilean-invariant momentum exchange with the full populations (delta-f + w rho0) */
        const double fin = fs_i + ww * rho0, fout = fnew + ww * rho0;
        const double Fx = (CX[ip] - uwx) * fin - (-CX[ip] - uwx) * fout;
        const double Fy = (CY[ip] - uwy) * fin - (-CY[ip] - uwy) * fout;
        const double Fz = (CZ[ip] - uwz) * fin - (-CZ[ip] - uwz) * fout;
        fHIp[3 * p] += Fx; fHIp[3 * p + 1] += Fy; fHIp[3 * p + 2] += Fz;
        torqp[3 * p] += ry * Fz - rz * Fy;
        torqp[3 * p + 1] += rz * Fx - rx * Fz;
        torqp[3 * p + 2] += rx * Fy - ry * Fx;
    }
}

/* ---- beads_filling: populations of nodes the particles uncovered ---------------------------------
 * own0: mask before the move, own: after.  ypglb/wp/omgp: the state AFTER the move of the particle
 * that owned the node.  Returns the number of refilled nodes (z,y,x order). */
static void feq19(double rho, double ux, double uy, double uz, double *fe)
{
    const double usqr = 1.5 * (ux * ux + uy * uy + uz * uz);
    fe[0] = (1.0 / 3.0) * (rho - usqr);
    for (int ip = 1; ip < NPOP; ++ip) {
        const double G = CX[ip] * ux + CY[ip] * uy + CZ[ip] * uz;
        const double ww = ip <= 6 ? 1.0 / 18.0 : 1.0 / 36.0;
        fe[ip] = ww * (rho + 3.0 * G + 4.5 * G * G - usqr);
    }
}

long po_refill(const po_geom *g, double *f, const int32_t *own0, const int32_t *own, const double *ypglb,
               const double *wp, const double *omgp)
{
    long nfill = 0;
    const size_t ntot = (size_t)g->nx * g->ny * g->nz;
    double *fnew = (double *)malloc(sizeof(double) * NPOP * (ntot ? 1 : 1) * 1);
    (void)fnew;
    free(fnew);
    /* two passes so that refilled nodes never feed each other: collect, then write */
    long cap = 1024, cnt = 0;
    size_t *where = (size_t *)malloc(sizeof(size_t) * (size_t)cap);
    double *vals = (double *)malloc(sizeof(double) * NPOP * (size_t)cap);
    for (int iz = 1; iz <= g->nz; ++iz)
        for (int iy = 1; iy <= g->ny; ++iy)
            for (int ix = 1; ix <= g->nx; ++ix) {
                const size_t n = nidx(g, ix, iy, iz);
                if (!(own0[n] > 0 && own[n] < 0)) continue;
                const int p = own0[n] - 1;
                double c[3] = {ypglb[3 * p], ypglb[3 * p + 1], ypglb[3 * p + 2]};
                const double xf[3] = {(double)ix - 0.5, (double)iy - 0.5, (double)iz - 0.5};
                if (c[1] - xf[1] > 0.5 * g->ny) c[1] -= g->ny;
                if (c[1] - xf[1] < -0.5 * g->ny) c[1] += g->ny;
                if (c[2] - xf[2] > 0.5 * g->nz) c[2] -= g->nz;
                if (c[2] - xf[2] < -0.5 * g->nz) c[2] += g->nz;
                const double rx = xf[0] - c[0], ry = xf[1] - c[1], rz = xf[2] - c[2];
                const double uwx = wp[3 * p] + (omgp[3 * p + 1] * rz - omgp[3 * p + 2] * ry);
                const double uwy = wp[3 * p + 1] + (omgp[3 * p + 2] * rx - omgp[3 * p] * rz);
                const double uwz = wp[3 * p + 2] + (omgp[3 * p] * ry - omgp[3 * p + 1] * rx);
                /* averaged density of the neighbours that were and are fluid; the one nearest the outward normal */
                double rsum = 0.0, best = -2.0;
                int nn = 0, jbest = 0;
                for (int ip = 1; ip < NPOP; ++ip) {
                    const int kx = ix + CX[ip], ky = wrap(iy + CY[ip], g->ny), kz = wrap(iz + CZ[ip], g->nz);
                    if (kx < 1 || kx > g->nx) continue;
                    const size_t m = nidx(g, kx, ky, kz);
                    if (own0[m] > 0 || own[m] > 0) continue;
                    double r = 0.0;
                    for (int k = 0; k < NPOP; ++k) r += f[k + NPOP * m];
                    rsum += r; ++nn;
                    const double cn = (CX[ip] * rx + CY[ip] * ry + CZ[ip] * rz) / sqrt((double)(CX[ip] * CX[ip] + CY[ip] * CY[ip] + CZ[ip] * CZ[ip]));
                    if (cn > best) { best = cn; jbest = ip; }
                }
                const double rbar = nn ? rsum / (double)nn : 0.0;
                double fe[NPOP], out[NPOP];
                feq19(rbar, uwx, uwy, uwz, fe);
                for (int k = 0; k < NPOP; ++k) out[k] = fe[k];
                if (jbest) {
                    const size_t m = nidx(g, ix + CX[jbest], wrap(iy + CY[jbest], g->ny), wrap(iz + CZ[jbest], g->nz));
                    const double *fm = f + NPOP * m;
                    double r = 0.0, jx = 0.0, jy = 0.0, jz = 0.0, fe2[NPOP];
                    for (int k = 0; k < NPOP; ++k) { r += fm[k]; jx += CX[k] * fm[k]; jy += CY[k] * fm[k]; jz += CZ[k] * fm[k]; }
                    feq19(r, jx, jy, jz, fe2);
                    for (int k = 0; k < NPOP; ++k) out[k] += fm[k] - fe2[k];
                }
                if (cnt == cap) {
                    cap *= 2;
                    where = (size_t *)realloc(where, sizeof(size_t) * (size_t)cap);
                    vals = (double *)realloc(vals, sizeof(double) * NPOP * (size_t)cap);
                }
                where[cnt] = n;
                memcpy(vals + NPOP * cnt, out, sizeof out);
                ++cnt; ++nfill;
            }
    for (long i = 0; i < cnt; ++i) memcpy(f + NPOP * where[i], vals + NPOP * i, sizeof(double) * NPOP);
    free(where); free(vals);
    return nfill;
}

/* ---- beads_lubforce: short-range repulsion (particle-particle, particle-wall) -------------------- */
void po_lubforce(const po_geom *g, int npart, const double *ypglb, double mingap, double mingap_w, double stf0,
                 double stf1, double stf0_w, double stf1_w, double fscale, double *flubp)
{
    memset(flubp, 0, sizeof(double) * 3 * (size_t)npart);
    const double R = g->rad;
    for (int i = 0; i < npart; ++i) {
        const double *a = ypglb + 3 * i;
        for (int j = 0; j < npart; ++j) {
            if (j == i) continue;
            const double *b = ypglb + 3 * j;
            double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
            if (dy > 0.5 * g->ny) dy -= g->ny;
            if (dy < -0.5 * g->ny) dy += g->ny;
            if (dz > 0.5 * g->nz) dz -= g->nz;
            if (dz < -0.5 * g->nz) dz += g->nz;
            const double d = sqrt(dx * dx + dy * dy + dz * dz);
            const double gap = d - 2.0 * R;
            if (gap >= mingap || d == 0.0) continue;
            double mag = fscale / stf0 * ((gap - mingap) / mingap) * ((gap - mingap) / mingap);
            if (gap < 0.0) mag += fscale / stf1 * (-gap / mingap);
            flubp[3 * i] += mag * dx / d; flubp[3 * i + 1] += mag * dy / d; flubp[3 * i + 2] += mag * dz / d;
        }
        /* channel walls at x = 0 and x = nx: image particle behind the wall */
        for (int s = 0; s < 2; ++s) {
            const double dxw = s == 0 ? a[0] : a[0] - (double)g->nx;          /* signed distance centre - wall */
            const double gap = fabs(dxw) - R;
            if (gap >= mingap_w) continue;
            double mag = fscale / stf0_w * ((gap - mingap_w) / mingap_w) * ((gap - mingap_w) / mingap_w);
            if (gap < 0.0) mag += fscale / stf1_w * (-gap / mingap_w);
            flubp[3 * i] += (dxw >= 0.0 ? mag : -mag);
        }
    }
}

/* ---- beads_move: Newton-Euler update with two-step averaged force ------------------------------- */
void po_move(const po_geom *g, int npart, double amp, double aip, const double *fHIp, const double *torqp,
             const double *flubp, double *forcepp, double *torqpp, const double *gforce, double *ypglb, double *wp,
             double *omgp, double *thetap)
{
    for (int p = 0; p < npart; ++p)
        for (int d = 0; d < 3; ++d) {
            const int k = 3 * p + d;
            const double F = 0.5 * (fHIp[k] + forcepp[k]) + flubp[k] + gforce[d];
            const double T = 0.5 * (torqp[k] + torqpp[k]);
            const double wnew = wp[k] + F / amp;
            const double onew = omgp[k] + T / aip;
            ypglb[k] += 0.5 * (wp[k] + wnew);
            thetap[k] += 0.5 * (omgp[k] + onew);
            wp[k] = wnew; omgp[k] = onew;
            forcepp[k] = fHIp[k]; torqpp[k] = torqp[k];
        }
    for (int p = 0; p < npart; ++p) {                      /* periodic in y and z */
        double *c = ypglb + 3 * p;
        if (c[1] >= (double)g->ny) c[1] -= g->ny;
        if (c[1] < 0.0) c[1] += g->ny;
        if (c[2] >= (double)g->nz) c[2] -= g->nz;
        if (c[2] < 0.0) c[2] += g->nz;
    }
}
