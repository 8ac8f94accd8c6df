// channel_driver.cpp -- the Channel-Flow driver in compiled code above the C-ABI.
//
// The reference's driver is Fortran (main.f90 + para.f90 + var_inc.f90 + initial.f90) and stays intact in a
// real deployment: it links collision_b200.f90, which forwards the seven hot-path subroutines to
// libd3q19b200.so.  This image has no Fortran compiler, so this file is the same caller in
// C++: it restates exactly the part of the driver that sits above the boundary --
//     para        para.f90:21-214     parameters, MRT constants                (struct VarInc, para())
//     allocarray  para.f90:418-503    the host arrays of module var_inc        (allocarray())
//     initvel     initial.f90:75-147  log-law / zero initial velocity          (initvel())
//     initpop     initial.f90:19-46   equilibrium populations                  (initpop())
//     main        main.f90:52-236     new-run sequence, pre-relaxation loop, time loop, diag cadence, probe
// -- and calls the hot path through the very d3q19_shim_* entry points the Fortran shim binds
// (include/d3q19_b200.h), with the host arrays bound once like the shim's d3q19_b200_ensure.  The host arrays
// are therefore exactly as current as the intact Fortran driver would find them (download policy of
// SURVEY.md 8(b)); `diag` and `probe` below read them on the host like saveload.f90 does.
//
// Ranks: the box has no MPI, so `--ranks N` runs the N ranks of the reference's job as N THREADS of this process, one
// GPU each (device = rank mod device count), z-slabs as para.f90:229-262 cuts them with nprocY = 1; what the reference
// does with MPI around the hot path (MPI_ALLREDUCE of the pre-relaxation error main.f90:80, the reductions and gathers
// of diag / outputuy / probe, the broadcast of the NCCL id the Fortran shim does with MPI_BCAST) goes through shared
// memory and a barrier here.  The library itself is used exactly as by one rank per process: one handle per rank.
//
//   channel_driver --nx 64 --ny 32 --nz 32 [--ranks 1] [--turbulent] [--nsteps 1000] [--ndiag 250] [--nflowout 100]
//                  [--prerelax [--prerelax-max 15000]] [--A9 0.0] [--scheme aa|ab|auto] [--strict]
//                  [--time-lmt 720 --time-buff 10 --ntime 10000] [--dump state.bin] [--dry-run]
// --dry-run stops after initpop (no GPU, no library call): used by the CPU tests to check para / initvel /
// initpop against the oracle bit for bit.  --dump writes nx,ny,nz,istep (int32) and f,rho,ux,uy,uz (fp64).
//
// Build (d3q19-single-phase_b200/build.py build_driver): g++ -std=c++17 -O2 -ffp-contract=off, so that every
// expression below is the IEEE evaluation of the Fortran source order.
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "d3q19_b200.h"

namespace {

struct VarInc {                     // the part of `module var_inc` the driver side touches
    int nx = 0, ny = 0, nz = 0, lx = 0, ly = 0, lz = 0, nxh = 0, npop = 19;
    int nproc = 1, myid = 0, indz = 0, globalz = 0;      // para.f90:219-262 with nprocY = 1: rank = indz
    int nsteps = 1000, istep0 = 0, istep = 0;            // para.f90:43-45
    int ndiag = 250, nflowout = 100, ntime = 10000;      // var_inc.f90:58-59
    bool laminar = true, ivel = false, ipart = false;    // para.f90:59,69,85,332
    int MRTtype = 2;
    double rho0 = 1.0, rhopart = 1.0, pi = 0, pi2 = 0;   // var_inc.f90:65,71
    double visc = 0, Rstar = 0, ustar = 0, force_in_y = 0, ystar = 0, force_mag = 1.0;
    double tau = 0, s1 = 0, s2 = 0, s4 = 0, s9 = 0, s10 = 0, s13 = 0, s16 = 0, omegepsl = 0, omegepslj = 0, omegxx = 0;
    double ww0 = 0, ww1 = 0, ww2 = 0;
    double rhoepsl = 1.0e-05;                            // para.f90:285
    int cix[19], ciy[19], ciz[19];
    std::vector<double> f, rho, rhop, ux, uy, uz, force_realx, force_realy, force_realz;
    std::vector<int32_t> ibnodes;                        // (0:lx+1,0:ly+1,0:lz+1), -1 = fluid
};

// para.f90:21-214 for run-time sizes (the reference fixes nx = nx7-1, ny = nz = nx7 at compile time)
void para(VarInc &v, int nx, int ny, int nz, bool laminar) {
    v.nx = nx; v.ny = ny; v.nz = nz;
    v.nxh = (nx + 1) / 2;                                // nx7/2 with nx = nx7-1, var_inc.f90:54
    v.pi = 4.0 * std::atan(1.0);
    v.pi2 = 2.0 * v.pi;
    v.laminar = laminar;
    if (!laminar) {                                      // para.f90:61-70
        v.visc = 0.0036;
        v.Rstar = 180.0;
        v.ustar = 2.0 * v.Rstar * v.visc / (double)nx;
        v.force_in_y = 2. * v.rho0 * v.ustar * v.ustar / (double)nx;
        v.ystar = v.visc / v.ustar;
        v.force_mag = 1.0;
        v.ivel = true;
        v.MRTtype = 1;
    } else {                                             // para.f90:76-86
        v.Rstar = 20;
        v.ustar = 0.05;
        v.visc = 2.0 * v.ustar * (double)nx / v.Rstar;
        v.force_in_y = 8.0 * v.visc * v.ustar / ((double)nx * (double)nx);
        v.ystar = v.visc / v.ustar;
        v.force_mag = 1.0;
        v.ivel = false;
        v.MRTtype = 2;
    }
}

// para.f90:106-143,172-206: everything that follows from visc and MRTtype
void para_mrt(VarInc &v) {
    v.tau = 3.0 * v.visc + 0.5;
    v.s9 = 1.0 / v.tau;
    v.s13 = v.s9;
    switch (v.MRTtype) {
    case 1: v.s1 = 1.5; v.s2 = 1.4; v.s4 = 1.2; v.s10 = 1.4; v.s16 = 1.98;
            v.omegepsl = 0.0; v.omegepslj = -475.0 / 63.0; v.omegxx = 0.0; break;
    case 2: v.s1 = v.s2 = v.s4 = v.s10 = v.s16 = v.s9;
            v.omegepsl = 3.0; v.omegepslj = -11.0 / 2.0; v.omegxx = -1.0 / 2.0; break;
    default: v.s1 = 1.8; v.s2 = v.s1; v.s4 = v.s9; v.s10 = v.s1; v.s16 = v.s1;
            v.omegepsl = 3.0; v.omegepslj = -11.0 / 2.0; v.omegxx = -1.0 / 2.0; break;
    }
    v.ww0 = 1.0 / 3.0; v.ww1 = 1.0 / 18.0; v.ww2 = 1.0 / 36.0;
    const int cx[19] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
    const int cy[19] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
    const int cz[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
    for (int i = 0; i < 19; ++i) { v.cix[i] = cx[i]; v.ciy[i] = cy[i]; v.ciz[i] = cz[i]; }
}

// the z-slab of this rank, para.f90:240-244 (uneven split) and :259-261 (global offset), nprocY = 1
void topology(VarInc &v, int nproc, int myid) {
    v.nproc = nproc; v.myid = myid; v.indz = myid;
    const int base = (v.nz - v.nz % nproc) / nproc, extra = v.nz - nproc * (v.nz / nproc);
    v.lx = v.nx; v.ly = v.ny;
    v.lz = v.indz < extra ? base + 1 : base;
    v.globalz = 0;
    for (int i = 0; i < v.indz; ++i) v.globalz += (i < extra ? base + 1 : base);
}

// allocarray, para.f90:418-503
void allocarray(VarInc &v) {
    const size_t n = (size_t)v.lx * v.ly * v.lz;
    v.f.assign(19 * n, 0.0);
    for (auto *a : {&v.rho, &v.rhop, &v.ux, &v.uy, &v.uz, &v.force_realx, &v.force_realy, &v.force_realz}) a->assign(n, 0.0);
    v.ibnodes.assign((size_t)(v.lx + 2) * (v.ly + 2) * (v.lz + 2), -1);      // para.f90:447
}

inline size_t at(const VarInc &v, int i, int j, int k) {          // (i,j,k) 1-based -> (lx,ly,lz) column major
    return (size_t)(i - 1) + (size_t)v.lx * ((size_t)(j - 1) + (size_t)v.ly * (size_t)(k - 1));
}

// initvel, initial.f90:75-147 (A9 is the reference's hard-wired 0.0 unless overridden)
void initvel(VarInc &v, double A9) {
    const double alpha = 1.0, beta9 = 1.0, cc = 60.0;
    const double ccc1 = -(double)v.ny / v.pi2 / alpha / v.ystar * A9 * v.ustar / cc / cc;
    std::fill(v.ux.begin(), v.ux.end(), 0.0);
    std::fill(v.uy.begin(), v.uy.end(), 0.0);
    std::fill(v.uz.begin(), v.uz.end(), 0.0);
    if (!v.ivel) return;                                           // goto 111
    for (int i = 1; i <= v.nxh; ++i) {                             // :104-115
        const double yplus = ((double)i - 0.5) / v.ystar;
        double u9;
        if (yplus < 10.8) u9 = yplus * v.ustar;
        else { u9 = std::log(yplus) / 0.41 + 5.0; u9 = u9 * v.ustar; }
        for (int k = 1; k <= v.lz; ++k)
            for (int j = 1; j <= v.ly; ++j) { v.uy[at(v, i, j, k)] = u9; v.uy[at(v, v.nx + 1 - i, j, k)] = u9; }
    }
    for (int k = 1; k <= v.lz; ++k) {                              // :119-144 (indy = 0)
        const int kk = k + v.indz * v.lz;                          // :120, the reference's offset (exact for even slabs)
        const double z9 = v.pi2 * ((double)kk - 0.5) / (double)v.nz;
        for (int j = 1; j <= v.ly; ++j) {
            const double y9 = v.pi2 * ((double)j - 0.5) / (double)v.ny;
            for (int i = 1; i <= v.nxh; ++i) {
                const double yplus = ((double)i - 0.5) / v.ystar;
                const double ccc9 = std::exp(-yplus / cc);
                double u9 = ccc1 * yplus * ccc9 * std::sin(alpha * y9 + beta9 * z9);
                v.uy[at(v, i, j, k)] = v.uy[at(v, i, j, k)] + u9;
                v.uy[at(v, v.nx + 1 - i, j, k)] = v.uy[at(v, v.nx + 1 - i, j, k)] + u9;
                const double ccc10 = A9 * v.ustar * (1. - ccc9 - yplus / cc * ccc9);
                u9 = ccc10 * std::cos(alpha * y9 + beta9 * z9);
                v.ux[at(v, i, j, k)] = v.ux[at(v, i, j, k)] + u9;
                v.ux[at(v, v.nx + 1 - i, j, k)] = v.ux[at(v, v.nx + 1 - i, j, k)] + u9;
            }
        }
    }
}

// initpop, initial.f90:19-46
void initpop(VarInc &v) {
    const size_t n = (size_t)v.lx * v.ly * v.lz;
    for (size_t m = 0; m < n; ++m) {
        double usqr = v.ux[m] * v.ux[m] + v.uy[m] * v.uy[m] + v.uz[m] * v.uz[m];
        usqr = 1.5 * usqr;
        v.rho[m] = 0.0;
        const double rho = v.rho[m];
        v.f[19 * m + 0] = v.ww0 * (rho - usqr);
        for (int ip = 1; ip < 19; ++ip) {
            const double G = ((double)v.cix[ip] * v.ux[m] + (double)v.ciy[ip] * v.uy[m] + (double)v.ciz[ip] * v.uz[m]);
            const double ww = ip <= 6 ? v.ww1 : v.ww2;
            v.f[19 * m + ip] = ww * (rho + 3.0 * G + 4.5 * G * G - usqr);
        }
    }
}

// ---- the seven subroutines: forwarded exactly like collision_b200.f90 does -----------------------------------
thread_local d3q19_handle *H = nullptr;          // one handle per rank (= per thread)

// what MPI is to the reference's driver: a barrier and a place where every rank can see every rank's numbers
struct World {
    int nranks = 1;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    unsigned gen = 0;
    std::vector<VarInc *> v;                      // every rank's module var_inc (read by others only between barriers)
    std::vector<std::vector<double>> slot;        // per-rank scratch of the reductions
    unsigned char nccl_id[128];
    int stop = 0;
    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned g = gen;
        if (++arrived == nranks) { arrived = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
    // MPI_ALLREDUCE(MAX) of one number
    double allreduce_max(int rank, double x) {
        slot[rank].assign(1, x);
        barrier();
        double m = slot[0][0];
        for (int r = 1; r < nranks; ++r) m = std::fmax(m, slot[r][0]);
        barrier();
        return m;
    }
} W;

void check(int rc, const char *what) {
    if (rc != 0) {
        std::fprintf(stderr, "channel_driver: %s failed: %s\n", what, d3q19_last_error());
        std::exit(2);                                              // the Fortran shim calls MPI_ABORT here
    }
}

void ensure(VarInc &v, int scheme, int math, int prerelax_max) {                     // d3q19_b200_ensure of the Fortran shim
    if (H) return;
    d3q19_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = D3Q19_ABI_VERSION;
    cfg.lx = v.lx; cfg.ly = v.ly; cfg.lz = v.lz;
    cfg.nx = v.nx; cfg.ny = v.ny; cfg.nz = v.nz;
    cfg.globalz = v.globalz; cfg.rank = v.indz; cfg.nranks = v.nproc;
    int32_t ndev = 1;
    check(d3q19_device_count(&ndev), "d3q19_device_count");
    cfg.device = v.myid % (ndev > 0 ? ndev : 1);
    if (v.nproc > 1) {                            // the Fortran shim: rank 0 makes the id, MPI_BCAST hands it out
        if (v.myid == 0) check(d3q19_nccl_unique_id(W.nccl_id), "d3q19_nccl_unique_id");
        W.barrier();
        std::memcpy(cfg.nccl_id, W.nccl_id, 128);
    }
    cfg.scheme = scheme; cfg.math = math; cfg.ipart = v.ipart ? 1 : 0; cfg.overlap = 1;
    cfg.s1 = v.s1; cfg.s2 = v.s2; cfg.s4 = v.s4; cfg.s9 = v.s9; cfg.s10 = v.s10; cfg.s13 = v.s13; cfg.s16 = v.s16;
    cfg.omegepsl = v.omegepsl; cfg.omegepslj = v.omegepslj; cfg.omegxx = v.omegxx;
    cfg.rhopart = v.rhopart;
    check(d3q19_create(&cfg, &H), "d3q19_create");
    check(d3q19_shim_bind_arrays(H, v.f.data(), v.rho.data(), v.ux.data(), v.uy.data(), v.uz.data(), v.force_realx.data(),
                                 v.force_realy.data(), v.force_realz.data(), v.ibnodes.data(), nullptr, 0, v.ndiag,
                                 v.nflowout, v.nsteps, v.istep0, v.ntime, prerelax_max, v.rhoepsl), "d3q19_shim_bind_arrays");
}
void FORCING(VarInc &v) { check(d3q19_shim_forcing(H, v.force_in_y, v.force_mag), "FORCING"); }
void rhoupdat() { check(d3q19_shim_rhoupdat(H), "rhoupdat"); }
void collision_MRT() { check(d3q19_shim_collision_mrt(H), "collision_MRT"); }
void macrovar(const VarInc &v) { check(d3q19_shim_macrovar(H, v.istep), "macrovar"); }

// ---- host post-processing the intact driver would do in saveload.f90, on the HOST arrays ------------------------
// diag, saveload.f90:1535-1640 (no particles): the numbers of one diag.dat line.  Every rank sums over its slab, the
// sums meet on rank 0 in rank order (MPI_REDUCE, :1544-1551), the largest speed is that of the first rank with a
// strictly larger value (:1649-1656) at its global location.
void diag(const VarInc &v) {
    const size_t n = (size_t)v.lx * v.ly * v.lz;
    double um = 0, vm = 0, wm = 0, ur = 0, vr = 0, wr = 0, vmax = 0, rhomax = -HUGE_VAL, rhomin = HUGE_VAL;
    int im = 0, jm = 0, km = 0;
    for (size_t m = 0; m < n; ++m) { um += v.ux[m]; vm += v.uy[m]; wm += v.uz[m]; }
    for (size_t m = 0; m < n; ++m) { ur += v.ux[m] * v.ux[m]; vr += v.uy[m] * v.uy[m]; wr += v.uz[m] * v.uz[m]; }
    for (int k = 1; k <= v.lz; ++k)
        for (int j = 1; j <= v.ly; ++j)
            for (int i = 1; i <= v.lx; ++i) {
                const size_t m = at(v, i, j, k);
                const double vel = std::sqrt(v.ux[m] * v.ux[m] + v.uy[m] * v.uy[m] + v.uz[m] * v.uz[m]);
                if (vel > vmax) { vmax = vel; im = i; jm = j; km = k + v.globalz; }
                if (v.rho[m] > rhomax) rhomax = v.rho[m];
                if (v.rho[m] < rhomin) rhomin = v.rho[m];
            }
    double nf = (double)n;
    if (W.nranks > 1) {
        W.slot[v.myid] = {um, vm, wm, ur, vr, wr, vmax, (double)im, (double)jm, (double)km, rhomax, rhomin, nf};
        W.barrier();
        if (v.myid == 0) {
            um = vm = wm = ur = vr = wr = nf = 0.0; vmax = 0.0; rhomax = -HUGE_VAL; rhomin = HUGE_VAL;
            for (int r = 0; r < W.nranks; ++r) {
                const std::vector<double> &q = W.slot[r];
                um += q[0]; vm += q[1]; wm += q[2]; ur += q[3]; vr += q[4]; wr += q[5]; nf += q[12];
                if (q[6] > vmax) { vmax = q[6]; im = (int)q[7]; jm = (int)q[8]; km = (int)q[9]; }
                rhomax = std::fmax(rhomax, q[10]); rhomin = std::fmin(rhomin, q[11]);
            }
        }
        W.barrier();
        if (v.myid != 0) return;
    }
    um /= nf; vm /= nf; wm /= nf;
    ur = std::sqrt(ur / nf - um * um); vr = std::sqrt(vr / nf - vm * vm); wr = std::sqrt(wr / nf - wm * wm);
    std::printf("diag %d %.16e %d %d %d %.16e %.16e %.16e %.16e %.16e %.16e %.16e %.16e %.16e\n", v.istep, vmax, im, jm, km,
                um / v.ustar, vm / v.ustar, wm / v.ustar, ur / v.ustar, vr / v.ustar, wr / v.ustar, 0.0, rhomax, rhomin);
}

// outputuy's profile columns, saveload.f90:917-935: plane mean of uy / ustar next to the analytic steady
// and start-up Poiseuille solutions the reference embeds (the reference gathers uy on rank 0, :861-887; here every
// rank sums its slab and rank 0 adds the sums in rank order)
void outputuy(const VarInc &v) {
    const double tstar = (double)v.istep * v.visc / (((double)v.nx / 2.0) * ((double)v.nx / 2.0));
    std::vector<double> sums;
    for (int i = 1; i <= v.lx; i += (v.lx > 16 ? v.lx / 8 : 1)) {
        double s = 0.0;
        for (int k = 1; k <= v.lz; ++k)
            for (int j = 1; j <= v.ly; ++j) s += v.uy[at(v, i, j, k)];
        sums.push_back(s);
    }
    if (W.nranks > 1) {
        W.slot[v.myid] = sums;
        W.barrier();
        if (v.myid == 0)
            for (int r = 1; r < W.nranks; ++r)
                for (size_t q = 0; q < sums.size(); ++q) sums[q] += W.slot[r][q];
        W.barrier();
        if (v.myid != 0) return;
    }
    std::printf("uy_profile %d", v.istep);
    size_t q = 0;
    for (int i = 1; i <= v.lx; i += (v.lx > 16 ? v.lx / 8 : 1), ++q) {
        const double xx0 = std::fabs((double)i - 0.5 - (double)v.nx / 2.0), xi = xx0 / ((double)v.nx / 2.0);
        double uut = 1.0 - xi * xi;
        for (int nn = 0; nn <= 25; ++nn) {
            const double a = ((double)nn + 0.5) * v.pi;
            uut -= 4.0 * ((nn % 2) ? -1.0 : 1.0) / (a * a * a) * std::exp(-a * a * tstar) * std::cos(a * xi);
        }
        std::printf(" %d:%.8e/%.8e", i, sums[q] / ((double)v.ly * v.nz) / v.ustar, v.laminar ? uut : 0.0);
    }
    std::printf("\n");
}

// probe, saveload.f90:4059-4100: the centre node of EVERY rank's local domain, gathered on rank 0 in rank order
void probe(const VarInc &v) {
    if (W.nranks > 1) W.barrier();                                 // MPI_BARRIER, :4069; the others' arrays are final
    if (v.myid != 0) { if (W.nranks > 1) W.barrier(); return; }
    for (int r = 0; r < W.nranks; ++r) {
        const VarInc &q = *W.v[r];
        const size_t m = at(q, q.lx / 2, q.ly / 2, q.lz / 2 > 0 ? q.lz / 2 : 1);
        if (r == 0) std::printf("probe %d %.16e %.16e %.16e\n", v.istep, q.ux[m], q.uy[m], q.uz[m]);
        else std::printf("probe_rank %d %d %.16e %.16e %.16e\n", r, v.istep, q.ux[m], q.uy[m], q.uz[m]);
    }
    if (W.nranks > 1) W.barrier();
}

// nx,ny,nz,istep, then f, rho, ux, uy, uz of the WHOLE channel: z is the slowest index, so a field is the ranks'
// slabs one after the other
void dump(const VarInc &v, const std::string &path) {
    if (W.nranks > 1) W.barrier();
    if (v.myid == 0) {
        FILE *fp = std::fopen(path.c_str(), "wb");
        if (!fp) { std::perror(path.c_str()); std::exit(3); }
        const int32_t hdr[4] = {v.nx, v.ny, v.nz, v.istep};
        std::fwrite(hdr, sizeof hdr, 1, fp);
        for (int a = 0; a < 5; ++a)
            for (int r = 0; r < W.nranks; ++r) {
                const VarInc &q = *W.v[r];
                const std::vector<double> *arr[5] = {&q.f, &q.rho, &q.ux, &q.uy, &q.uz};
                std::fwrite(arr[a]->data(), sizeof(double), arr[a]->size(), fp);
            }
        std::fclose(fp);
    }
    if (W.nranks > 1) W.barrier();
}

struct Options {
    int nx = 64, ny = 32, nz = 32, nsteps = 1000, ndiag = 250, nflowout = 100, prerelax_max = 15000, ranks = 1;
    int scheme = D3Q19_SCHEME_AUTO, math = D3Q19_MATH_FAST, mrttype = 0;
    bool laminar = true, prerelax = false, dry = false;
    double A9 = 0.0, ustar_over = 0.0;
    double time_lmt = 720.0, time_buff = 10.0;          // wall-clock limit and save buffer in minutes (para.f90:50-52)
    int ntime = 10000;                                  // var_inc.f90:59
    std::string dump_path;
};

// main.f90:18-236 as one rank executes it
void rank_main(const Options &o, int myid, VarInc &v) {
    const bool root = myid == 0;
    para(v, o.nx, o.ny, o.nz, o.laminar);                           // main.f90:41
    if (o.ustar_over > 0.0) {                                       // wall units of another channel (para.f90:64-66)
        v.ustar = o.ustar_over;
        v.force_in_y = 2. * v.rho0 * v.ustar * v.ustar / (double)o.nx;
        v.ystar = v.visc / v.ustar;
    }
    if (o.mrttype) v.MRTtype = o.mrttype;
    para_mrt(v);
    v.nsteps = o.nsteps; v.ndiag = o.ndiag; v.nflowout = o.nflowout; v.ntime = o.ntime;
    const double time_bond = (o.time_lmt - o.time_buff) * 60.0;     // para.f90:54
    topology(v, o.ranks, myid);                                     // para.f90:219-262
    allocarray(v);                                                  // main.f90:44
    if (root)
        std::printf("para nx %d ny %d nz %d visc %.17g ustar %.17g force_in_y %.17g ystar %.17g tau %.17g MRTtype %d\n", v.nx,
                    v.ny, v.nz, v.visc, v.ustar, v.force_in_y, v.ystar, v.tau, v.MRTtype);

    initvel(v, o.A9);                                               // main.f90:58
    if (o.dry) {
        initpop(v);
        if (!o.dump_path.empty()) dump(v, o.dump_path);
        return;
    }
    ensure(v, o.scheme, o.math, o.prerelax_max);
    FORCING(v);                                                     // main.f90:61
    initpop(v);                                                     // main.f90:65
    check(d3q19_shim_sync_f_to_device(H), "host f changed");        // what the shim does after initpop / loadcntdflow
    v.istep = 0;
    if (o.prerelax) {                                               // main.f90:70-90
        for (;;) {
            v.rhop = v.rho;
            rhoupdat();
            collision_MRT();
            double rhoerr = 0.0;
            for (size_t m = 0; m < v.rho.size(); ++m) rhoerr = std::fmax(rhoerr, std::fabs(v.rho[m] - v.rhop[m]));
            if (W.nranks > 1) rhoerr = W.allreduce_max(myid, rhoerr);                       // :80
            if (root) std::printf("prerelax %d %.16e\n", v.istep, rhoerr);
            if (rhoerr <= v.rhoepsl || v.istep > o.prerelax_max) {
                if (root) std::printf("final relaxation => %d %.16e\n", v.istep, rhoerr);
                break;
            }
            v.istep = v.istep + 1;
        }
        // saveinitflow (main.f90:101) would write the host f here: the shim made it current in the last iteration
    }
    const auto t_up0 = std::chrono::steady_clock::now();
    macrovar(v);                                                    // main.f90:102 (uploads f if nothing has yet)
    if (root)
        std::printf("first macrovar (incl. upload of f when it is the first device call) %.3f s\n",
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - t_up0).count());
    v.istep0 = 0;
    v.istep = v.istep0;
    FORCING(v);                                                     // main.f90:132
    macrovar(v);                                                    // main.f90:136
    if (W.nranks > 1) W.barrier();
    const auto t_loop0 = std::chrono::steady_clock::now();          // time_start = MPI_WTIME(), main.f90:137
    int stopped_at = 0;
    for (v.istep = v.istep0 + 1; v.istep <= v.istep0 + v.nsteps; ++v.istep) {       // main.f90:142-208
        collision_MRT();                                            // :157
        macrovar(v);                                                // :161
        if (v.ndiag > 0 && v.istep % v.ndiag == 0) diag(v);         // :171
        if (v.nflowout > 0 && v.istep % v.nflowout == 0) outputuy(v);   // :184 -> saveload.f90:696,848
        if (v.ntime > 0 && v.istep % v.ntime == 0) {                // :197-206: leave the loop when the wall-clock budget is spent
            double time_max = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_loop0).count();
            if (W.nranks > 1) time_max = W.allreduce_max(myid, time_max);                   // :203
            if (time_max > time_bond) {                             // (the shim made rho,u current on this step for probe)
                if (root) std::printf("time budget: %.1f s > %.1f s, leaving the loop after step %d\n", time_max, time_bond, v.istep);
                stopped_at = v.istep;
                break;
            }
        }
    }
    v.istep = stopped_at ? stopped_at : v.istep0 + v.nsteps;
    check(d3q19_sync(H), "d3q19_sync");
    if (W.nranks > 1) W.barrier();                                  // main.f90:217
    if (root) {
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_loop0).count();   // main.f90:214-218
        std::printf("time loop %.3f s, %.1f MLUPS\n", dt, (double)v.nx * v.ny * v.nz * (v.istep - v.istep0) / dt / 1e6);
    }
    probe(v);                                                       // main.f90:221
    if (!o.dump_path.empty()) {
        check(d3q19_shim_sync_f_to_host(H), "sync f to host");      // what savecntdflow needs (saveload.f90:227)
        dump(v, o.dump_path);
    }
    if (W.nranks > 1) W.barrier();                                  // nobody frees arrays a neighbour may still store into
    check(d3q19_destroy(H), "d3q19_destroy");
    H = nullptr;
}

}  // namespace

int main(int argc, char **argv) {
    Options o;
    for (int a = 1; a < argc; ++a) {
        const std::string s = argv[a];
        auto next = [&]() -> const char * { if (a + 1 >= argc) { std::fprintf(stderr, "missing value after %s\n", s.c_str()); std::exit(1); } return argv[++a]; };
        if (s == "--nx") o.nx = std::atoi(next());
        else if (s == "--ny") o.ny = std::atoi(next());
        else if (s == "--nz") o.nz = std::atoi(next());
        else if (s == "--ranks") o.ranks = std::atoi(next());
        else if (s == "--nsteps") o.nsteps = std::atoi(next());
        else if (s == "--ndiag") o.ndiag = std::atoi(next());
        else if (s == "--nflowout") o.nflowout = std::atoi(next());
        else if (s == "--turbulent") o.laminar = false;
        else if (s == "--laminar") o.laminar = true;
        else if (s == "--prerelax") o.prerelax = true;
        else if (s == "--prerelax-max") o.prerelax_max = std::atoi(next());
        else if (s == "--A9") o.A9 = std::atof(next());
        else if (s == "--ustar") o.ustar_over = std::atof(next());
        else if (s == "--mrttype") o.mrttype = std::atoi(next());
        else if (s == "--strict") o.math = D3Q19_MATH_STRICT;
        else if (s == "--scheme") { const std::string t = next(); o.scheme = t == "aa" ? D3Q19_SCHEME_AA : (t == "ab" ? D3Q19_SCHEME_AB : D3Q19_SCHEME_AUTO); }
        else if (s == "--time-lmt") o.time_lmt = std::atof(next());
        else if (s == "--time-buff") o.time_buff = std::atof(next());
        else if (s == "--ntime") o.ntime = std::atoi(next());
        else if (s == "--dump") o.dump_path = next();
        else if (s == "--dry-run") o.dry = true;
        else { std::fprintf(stderr, "unknown option %s (see the header of channel_driver.cpp)\n", s.c_str()); return 1; }
    }
    if (o.ranks < 1 || o.ranks > o.nz) { std::fprintf(stderr, "--ranks must be between 1 and nz\n"); return 1; }
    W.nranks = o.ranks;
    W.slot.resize(o.ranks);
    std::vector<VarInc> vars(o.ranks);
    for (int r = 0; r < o.ranks; ++r) W.v.push_back(&vars[r]);
    if (o.ranks == 1) {
        rank_main(o, 0, vars[0]);
        return 0;
    }
    std::vector<std::thread> th;
    for (int r = 0; r < o.ranks; ++r) th.emplace_back([&, r] { rank_main(o, r, vars[r]); });
    for (auto &t : th) t.join();
    return 0;
}
