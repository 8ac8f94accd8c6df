// Host-compiled unit-test harness for d3q19-single-phase_b200/csrc/collide.cuh.
// Test infrastructure only: built by tests/test_collide_host.py with g++ (-ffp-contract=off)
// so the node-local algebra of the CUDA kernels can be checked on the GPU-less build box.
#include "collide.cuh"

extern "C" {

// mode 0: strict with own moments (macrovar order); 1: strict with given moments;
// mode 2: fast own moments; 3: fast given moments.  f is [n][19], arrays are [n].
void collide_nodes(int mode, long n, double *f, const double *rho, const double *ux, const double *uy,
                   const double *uz, double Fx, double Fy, double Fz, double shift, const double *mrt10) {
    d3q::Mrt c{mrt10[0], mrt10[1], mrt10[2], mrt10[3], mrt10[4], mrt10[5], mrt10[6], mrt10[7], mrt10[8], mrt10[9]};
    for (long k = 0; k < n; ++k) {
        double(&fk)[19] = *reinterpret_cast<double(*)[19]>(f + 19 * k);
        if (mode == 0) {
            double r, a, b, cc;
            d3q::moments_strict(fk, Fx, Fy, Fz, r, a, b, cc);
            d3q::collide_strict(fk, r - shift, a, b, cc, Fx, Fy, Fz, c);
        } else if (mode == 1) {
            d3q::collide_strict(fk, rho[k], ux[k], uy[k], uz[k], Fx, Fy, Fz, c);
        } else if (mode == 2) {
            d3q::collide_fast<true>(fk, 0, 0, 0, 0, Fx, Fy, Fz, shift, c);
        } else {
            d3q::collide_fast<false>(fk, rho[k], ux[k], uy[k], uz[k], Fx, Fy, Fz, 0.0, c);
        }
    }
}

void moments_nodes(long n, const double *f, double Fx, double Fy, double Fz, double *rho, double *ux, double *uy,
                   double *uz) {
    for (long k = 0; k < n; ++k) {
        const double(&fk)[19] = *reinterpret_cast<const double(*)[19]>(f + 19 * k);
        d3q::moments_strict(fk, Fx, Fy, Fz, rho[k], ux[k], uy[k], uz[k]);
    }
}
}
