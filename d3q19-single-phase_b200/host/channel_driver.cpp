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
// One rank (nproc = 1): the box has no MPI; the multi-GPU path is driven from bench.py / tests/mgpu_worker.py.
//
//   channel_driver --nx 64 --ny 32 --nz 32 [--turbulent] [--nsteps 1000] [--ndiag 250] [--nflowout 100]
//                  [--prerelax [--prerelax-max 15000]] [--A9 0.0] [--scheme aa|ab|auto] [--strict]
//                  [--dump state.bin] [--dry-run]
// --dry-run stops after initpop (no GPU, no library call): used by the CPU tests to check para / initvel /
// initpop against the oracle bit for bit.  --dump writes nx,ny,nz,istep (int32) and f,rho,ux,uy,uz (fp64).
//
// Build (d3q19-single-phase_b200/build.py build_driver): g++ -std=c++17 -O2 -ffp-contract=off, so that every
// expression below is the IEEE evaluation of the Fortran source order.
#include <chrono>
#include <cmath>
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

// allocarray, para.f90:418-503 (one rank: lx,ly,lz = nx,ny,nz)
void allocarray(VarInc &v) {
    v.lx = v.nx; v.ly = v.ny; v.lz = v.nz;
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
    for (int k = 1; k <= v.lz; ++k) {                              // :119-144 (indy = indz = 0)
        const double z9 = v.pi2 * ((double)k - 0.5) / (double)v.nz;
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
d3q19_handle *H = nullptr;

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
    cfg.globalz = 0; cfg.rank = 0; cfg.nranks = 1; cfg.device = 0;
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
// diag, saveload.f90:1535-1640 (one rank, no particles): the numbers of one diag.dat line
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
                if (vel > vmax) { vmax = vel; im = i; jm = j; km = k; }
                if (v.rho[m] > rhomax) rhomax = v.rho[m];
                if (v.rho[m] < rhomin) rhomin = v.rho[m];
            }
    const double nf = (double)n;
    um /= nf; vm /= nf; wm /= nf;
    ur = std::sqrt(ur / nf - um * um); vr = std::sqrt(vr / nf - vm * vm); wr = std::sqrt(wr / nf - wm * wm);
    std::printf("diag %d %.16e %d %d %d %.16e %.16e %.16e %.16e %.16e %.16e %.16e %.16e %.16e\n", v.istep, vmax, im, jm, km,
                um / v.ustar, vm / v.ustar, wm / v.ustar, ur / v.ustar, vr / v.ustar, wr / v.ustar, 0.0, rhomax, rhomin);
}

// outputuy's profile columns, saveload.f90:917-935: plane mean of uy / ustar next to the analytic steady
// and start-up Poiseuille solutions the reference embeds
void outputuy(const VarInc &v) {
    const double tstar = (double)v.istep * v.visc / (((double)v.nx / 2.0) * ((double)v.nx / 2.0));
    std::printf("uy_profile %d", v.istep);
    for (int i = 1; i <= v.lx; i += (v.lx > 16 ? v.lx / 8 : 1)) {
        double s = 0.0;
        for (int k = 1; k <= v.lz; ++k)
            for (int j = 1; j <= v.ly; ++j) s += v.uy[at(v, i, j, k)];
        const double xx0 = std::fabs((double)i - 0.5 - (double)v.nx / 2.0), xi = xx0 / ((double)v.nx / 2.0);
        double uut = 1.0 - xi * xi;
        for (int nn = 0; nn <= 25; ++nn) {
            const double a = ((double)nn + 0.5) * v.pi;
            uut -= 4.0 * ((nn % 2) ? -1.0 : 1.0) / (a * a * a) * std::exp(-a * a * tstar) * std::cos(a * xi);
        }
        std::printf(" %d:%.8e/%.8e", i, s / ((double)v.ly * v.lz) / v.ustar, v.laminar ? uut : 0.0);
    }
    std::printf("\n");
}

// probe, saveload.f90:4059-4100: the centre node's velocity
void probe(const VarInc &v) {
    const size_t m = at(v, v.lx / 2, v.ly / 2, v.lz / 2);
    std::printf("probe %d %.16e %.16e %.16e\n", v.istep, v.ux[m], v.uy[m], v.uz[m]);
}

void dump(const VarInc &v, const std::string &path) {
    FILE *fp = std::fopen(path.c_str(), "wb");
    if (!fp) { std::perror(path.c_str()); std::exit(3); }
    const int32_t hdr[4] = {v.nx, v.ny, v.nz, v.istep};
    std::fwrite(hdr, sizeof hdr, 1, fp);
    for (const auto *a : {&v.f, &v.rho, &v.ux, &v.uy, &v.uz}) std::fwrite(a->data(), sizeof(double), a->size(), fp);
    std::fclose(fp);
}

}  // namespace

int main(int argc, char **argv) {
    int nx = 64, ny = 32, nz = 32, nsteps = 1000, ndiag = 250, nflowout = 100, prerelax_max = 15000;
    int scheme = D3Q19_SCHEME_AUTO, math = D3Q19_MATH_FAST, mrttype = 0;
    bool laminar = true, prerelax = false, dry = false;
    double A9 = 0.0, ustar_over = 0.0;
    double time_lmt = 720.0, time_buff = 10.0;          // wall-clock limit and save buffer in minutes (para.f90:50-52)
    int ntime = 10000;                                  // var_inc.f90:59
    std::string dump_path;
    for (int a = 1; a < argc; ++a) {
        const std::string s = argv[a];
        auto next = [&]() -> const char * { if (a + 1 >= argc) { std::fprintf(stderr, "missing value after %s\n", s.c_str()); std::exit(1); } return argv[++a]; };
        if (s == "--nx") nx = std::atoi(next());
        else if (s == "--ny") ny = std::atoi(next());
        else if (s == "--nz") nz = std::atoi(next());
        else if (s == "--nsteps") nsteps = std::atoi(next());
        else if (s == "--ndiag") ndiag = std::atoi(next());
        else if (s == "--nflowout") nflowout = std::atoi(next());
        else if (s == "--turbulent") laminar = false;
        else if (s == "--laminar") laminar = true;
        else if (s == "--prerelax") prerelax = true;
        else if (s == "--prerelax-max") prerelax_max = std::atoi(next());
        else if (s == "--A9") A9 = std::atof(next());
        else if (s == "--ustar") ustar_over = std::atof(next());
        else if (s == "--mrttype") mrttype = std::atoi(next());
        else if (s == "--strict") math = D3Q19_MATH_STRICT;
        else if (s == "--scheme") { const std::string t = next(); scheme = t == "aa" ? D3Q19_SCHEME_AA : (t == "ab" ? D3Q19_SCHEME_AB : D3Q19_SCHEME_AUTO); }
        else if (s == "--time-lmt") time_lmt = std::atof(next());
        else if (s == "--time-buff") time_buff = std::atof(next());
        else if (s == "--ntime") ntime = std::atoi(next());
        else if (s == "--dump") dump_path = next();
        else if (s == "--dry-run") dry = true;
        else { std::fprintf(stderr, "unknown option %s (see the header of channel_driver.cpp)\n", s.c_str()); return 1; }
    }
    VarInc v;
    para(v, nx, ny, nz, laminar);                                   // main.f90:41
    if (ustar_over > 0.0) {                                         // wall units of another channel (para.f90:64-66)
        v.ustar = ustar_over;
        v.force_in_y = 2. * v.rho0 * v.ustar * v.ustar / (double)nx;
        v.ystar = v.visc / v.ustar;
    }
    if (mrttype) v.MRTtype = mrttype;
    para_mrt(v);
    v.nsteps = nsteps; v.ndiag = ndiag; v.nflowout = nflowout; v.ntime = ntime;
    const double time_bond = (time_lmt - time_buff) * 60.0;        // para.f90:54
    allocarray(v);                                                  // main.f90:44
    std::printf("para nx %d ny %d nz %d visc %.17g ustar %.17g force_in_y %.17g ystar %.17g tau %.17g MRTtype %d\n", v.nx, v.ny,
                v.nz, v.visc, v.ustar, v.force_in_y, v.ystar, v.tau, v.MRTtype);

    initvel(v, A9);                                                 // main.f90:58
    if (dry) {
        initpop(v);
        if (!dump_path.empty()) dump(v, dump_path);
        return 0;
    }
    ensure(v, scheme, math, prerelax_max);
    FORCING(v);                                                     // main.f90:61
    initpop(v);                                                     // main.f90:65
    check(d3q19_shim_sync_f_to_device(H), "host f changed");        // what the shim does after initpop / loadcntdflow
    v.istep = 0;
    if (prerelax) {                                                 // main.f90:70-90
        for (;;) {
            v.rhop = v.rho;
            rhoupdat();
            collision_MRT();
            double rhoerr = 0.0;
            for (size_t m = 0; m < v.rho.size(); ++m) rhoerr = std::fmax(rhoerr, std::fabs(v.rho[m] - v.rhop[m]));
            std::printf("prerelax %d %.16e\n", v.istep, rhoerr);
            if (rhoerr <= v.rhoepsl || v.istep > prerelax_max) {
                std::printf("final relaxation => %d %.16e\n", v.istep, rhoerr);
                break;
            }
            v.istep = v.istep + 1;
        }
        // saveinitflow (main.f90:101) would write the host f here: the shim made it current in the last iteration
    }
    const auto t_up0 = std::chrono::steady_clock::now();
    macrovar(v);                                                    // main.f90:102 (uploads f if nothing has yet)
    std::printf("first macrovar (incl. upload of f when it is the first device call) %.3f s\n",
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_up0).count());
    v.istep0 = 0;
    v.istep = v.istep0;
    FORCING(v);                                                     // main.f90:132
    macrovar(v);                                                    // main.f90:136
    const auto t_loop0 = std::chrono::steady_clock::now();          // time_start = MPI_WTIME(), main.f90:137
    int stopped_at = 0;
    for (v.istep = v.istep0 + 1; v.istep <= v.istep0 + v.nsteps; ++v.istep) {       // main.f90:142-208
        collision_MRT();                                            // :157
        macrovar(v);                                                // :161
        if (v.ndiag > 0 && v.istep % v.ndiag == 0) diag(v);         // :171
        if (v.nflowout > 0 && v.istep % v.nflowout == 0) outputuy(v);   // :184 -> saveload.f90:696,848
        if (v.ntime > 0 && v.istep % v.ntime == 0) {                // :197-206: leave the loop when the wall-clock budget is spent
            const double time_max = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_loop0).count();
            if (time_max > time_bond) {                             // (the shim made rho,u current on this step for probe)
                std::printf("time budget: %.1f s > %.1f s, leaving the loop after step %d\n", time_max, time_bond, v.istep);
                stopped_at = v.istep;
                break;
            }
        }
    }
    v.istep = stopped_at ? stopped_at : v.istep0 + v.nsteps;
    check(d3q19_sync(H), "d3q19_sync");
    {
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_loop0).count();   // main.f90:214-218
        std::printf("time loop %.3f s, %.1f MLUPS\n", dt, (double)v.nx * v.ny * v.nz * (v.istep - v.istep0) / dt / 1e6);
    }
    probe(v);                                                       // main.f90:221
    if (!dump_path.empty()) {
        check(d3q19_shim_sync_f_to_host(H), "sync f to host");      // what savecntdflow needs (saveload.f90:227)
        dump(v, dump_path);
    }
    check(d3q19_destroy(H), "d3q19_destroy");
    return 0;
}
