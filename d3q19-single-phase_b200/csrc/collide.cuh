// collide.cuh -- the node-local part of collision_MRT (collision.f90:54-220 and the
// "+ 0.5*Fbar" of :230) and of macrovar (collision.f90:396-418), in two arithmetics:
//
//   strict  the reference's expressions in the reference's order, every + - * rounded
//           separately (no FMA contraction): bit-identical to the CPU oracle.
//   fast    the same map written as  f* = f + Minv * delta  with
//           delta_k = mF_k - s_k (m_k(f) + mF_k/2 - meq_k)   (non-conserved moments)
//           using opposite-pair sums/differences; ~235 flops instead of ~560, FMA allowed.
//           Agrees with strict to rounding (tests/test_collide_host.py, < 1e-14).
//
// Both are __host__ __device__ so the algebra is unit-tested with g++ on the build box;
// the product only ever calls them from the CUDA kernels in kernels.cuh.
#pragma once
#include "lattice.cuh"

namespace d3q {

// ---- a double whose + - * are never contracted ------------------------------------------
struct R {
    double v;
    D3Q_HD R() : v(0.0) {}
    D3Q_HD R(double x) : v(x) {}
};
D3Q_HD R operator+(R a, R b) {
#ifdef __CUDA_ARCH__
    return R(__dadd_rn(a.v, b.v));
#else
    return R(a.v + b.v);
#endif
}
D3Q_HD R operator-(R a, R b) {
#ifdef __CUDA_ARCH__
    return R(__dsub_rn(a.v, b.v));
#else
    return R(a.v - b.v);
#endif
}
D3Q_HD R operator*(R a, R b) {
#ifdef __CUDA_ARCH__
    return R(__dmul_rn(a.v, b.v));
#else
    return R(a.v * b.v);
#endif
}
D3Q_HD R operator-(R a) { return R(-a.v); }

// ---- moments in macrovar's order (collision.f90:398-417) ---------------------------------
// rho9 = f0+f1+...+f18 left to right; u = (c-weighted sums) + F/2.
D3Q_HD void moments_strict(const double (&f)[NPOP], double fx, double fy, double fz,
                           double &rho, double &ux, double &uy, double &uz) {
    R sum1 = R(f[7]) - R(f[10]);
    R sum2 = R(f[9]) - R(f[8]);
    R sum3 = R(f[11]) - R(f[14]);
    R sum4 = R(f[13]) - R(f[12]);
    R sum5 = R(f[15]) - R(f[18]);
    R sum6 = R(f[17]) - R(f[16]);
    R ux9 = R(f[1]) - R(f[2]) + sum1 + sum2 + sum3 + sum4;
    R uy9 = R(f[3]) - R(f[4]) + sum1 - sum2 + sum5 + sum6;
    R uz9 = R(f[5]) - R(f[6]) + sum3 - sum4 + sum5 - sum6;
    R rho9 = R(f[0]);
#pragma unroll
    for (int ip = 1; ip < NPOP; ++ip) rho9 = rho9 + R(f[ip]);
    ux = (ux9 + R(fx * 0.5)).v;   // force_realx/2. : exact scaling
    uy = (uy9 + R(fy * 0.5)).v;
    uz = (uz9 + R(fz * 0.5)).v;
    rho = rho9.v;
}

// ---- strict collision: collision.f90:56-220 + the final "+ 0.5*Fbar" ---------------------
D3Q_HD void collide_strict(double (&f)[NPOP], double rho_in, double ux_in, double uy_in, double uz_in,
                           double fx_in, double fy_in, double fz_in, const Mrt &c) {
    const R coef1(-2.0 / 3.0), coef2(-11.0), coef3(8.0), coef4(-4.0), coef5(2.0);
    const R coef3i(1.0 / 8.0), coef4i(1.0 / -4.0);
    const R val8(12.0);
    const R val1i(1.0 / 19.0), val2i(1.0 / 2394.0), val3i(1.0 / 252.0), val4i(1.0 / 10.0), val5i(1.0 / 40.0),
        val6i(1.0 / 36.0), val7i(1.0 / 72.0), val8i(1.0 / 12.0), val9i(1.0 / 24.0);
    const R ww1(1.0 / 18.0), ww2(1.0 / 36.0);
    const R s1(c.s1), s2(c.s2), s4(c.s4), s9(c.s9), s10(c.s10), s13(c.s13), s16(c.s16);
    const R omegepsl(c.omegepsl), omegepslj(c.omegepslj), omegxx(c.omegxx);

    const R rho9(rho_in), ux9(ux_in), uy9(uy_in), uz9(uz_in);
    const R ux9s = ux9 * ux9, uy9s = uy9 * uy9, uz9s = uz9 * uz9;
    const R fx9(fx_in), fy9(fy_in), fz9(fz_in);
    const R G3 = ux9 * fx9 + uy9 * fy9 + uz9 * fz9;

    R Fbar[NPOP], f9[NPOP];
    Fbar[0] = -G3;
    static_for<NPOP - 1>([&](auto ic) {
        constexpr int ip = decltype(ic)::value + 1;
        const R G1 = R(double(dir_cx(ip))) * fx9 + R(double(dir_cy(ip))) * fy9 + R(double(dir_cz(ip))) * fz9;
        const R G2 = R(double(dir_cx(ip))) * ux9 + R(double(dir_cy(ip))) * uy9 + R(double(dir_cz(ip))) * uz9;
        const R ww = ip <= 6 ? ww1 : ww2;
        Fbar[ip] = ww * (R(3.) * G1 + R(9.) * G1 * G2 - R(3.) * G3);
    });
#pragma unroll
    for (int ip = 0; ip < NPOP; ++ip) f9[ip] = R(f[ip]) + R(0.5) * Fbar[ip];

    const R t1 = ux9s + uy9s + uz9s;
    const R eqm1 = R(-11.0) * rho9 + R(19.0) * t1;
    const R eqm2 = omegepsl * rho9 + omegepslj * t1;
    const R eqm3 = coef1 * ux9;
    const R eqm4 = coef1 * uy9;
    const R eqm5 = coef1 * uz9;
    const R eqm6 = R(2.0) * ux9s - uy9s - uz9s;
    const R eqm7 = omegxx * eqm6;
    const R eqm8 = uy9s - uz9s;
    const R eqm9 = omegxx * eqm8;
    const R eqm10 = ux9 * uy9;
    const R eqm11 = uy9 * uz9;
    const R eqm12 = ux9 * uz9;
    const R eqm13(0.0), eqm14(0.0), eqm15(0.0);

    const R sum1 = f9[1] + f9[2] + f9[3] + f9[4] + f9[5] + f9[6];
    const R sum2 = f9[7] + f9[8] + f9[9] + f9[10] + f9[11] + f9[12] + f9[13] + f9[14] + f9[15] + f9[16] + f9[17] + f9[18];
    const R sum3 = f9[7] - f9[8] + f9[9] - f9[10] + f9[11] - f9[12] + f9[13] - f9[14];
    const R sum4 = f9[7] + f9[8] - f9[9] - f9[10] + f9[15] - f9[16] + f9[17] - f9[18];
    const R sum5 = f9[11] + f9[12] - f9[13] - f9[14] + f9[15] + f9[16] - f9[17] - f9[18];
    const R sum6 = f9[1] + f9[2];
    const R sum7 = f9[3] + f9[4] + f9[5] + f9[6];
    const R sum8 = f9[7] + f9[8] + f9[9] + f9[10] + f9[11] + f9[12] + f9[13] + f9[14];
    const R sum9 = f9[15] + f9[16] + f9[17] + f9[18];
    const R sum10 = f9[3] + f9[4] - f9[5] - f9[6];
    const R sum11 = f9[7] + f9[8] + f9[9] + f9[10] - f9[11] - f9[12] - f9[13] - f9[14];

    const R evlm1 = R(-30.0) * f9[0] + coef2 * sum1 + coef3 * sum2;
    const R evlm2 = R(12.0) * f9[0] + coef4 * sum1 + sum2;
    const R evlm3 = coef4 * (f9[1] - f9[2]) + sum3;
    const R evlm4 = coef4 * (f9[3] - f9[4]) + sum4;
    const R evlm5 = coef4 * (f9[5] - f9[6]) + sum5;
    const R evlm6 = coef5 * sum6 - sum7 + sum8 - coef5 * sum9;
    const R evlm7 = coef4 * sum6 + coef5 * sum7 + sum8 - coef5 * sum9;
    const R evlm8 = sum10 + sum11;
    const R evlm9 = -(coef5 * sum10) + sum11;
    const R evlm10 = f9[7] - f9[8] - f9[9] + f9[10];
    const R evlm11 = f9[15] - f9[16] - f9[17] + f9[18];
    const R evlm12 = f9[11] - f9[12] - f9[13] + f9[14];
    const R evlm13 = f9[7] - f9[8] + f9[9] - f9[10] - f9[11] + f9[12] - f9[13] + f9[14];
    const R evlm14 = -f9[7] - f9[8] + f9[9] + f9[10] + f9[15] - f9[16] + f9[17] - f9[18];
    const R evlm15 = f9[11] + f9[12] - f9[13] - f9[14] - f9[15] - f9[16] + f9[17] + f9[18];

    const R eqmc1 = evlm1 - s1 * (evlm1 - eqm1);
    const R eqmc2 = evlm2 - s2 * (evlm2 - eqm2);
    const R eqmc3 = evlm3 - s4 * (evlm3 - eqm3);
    const R eqmc4 = evlm4 - s4 * (evlm4 - eqm4);
    const R eqmc5 = evlm5 - s4 * (evlm5 - eqm5);
    const R eqmc6 = evlm6 - s9 * (evlm6 - eqm6);
    const R eqmc7 = evlm7 - s10 * (evlm7 - eqm7);
    const R eqmc8 = evlm8 - s9 * (evlm8 - eqm8);
    const R eqmc9 = evlm9 - s10 * (evlm9 - eqm9);
    const R eqmc10 = evlm10 - s13 * (evlm10 - eqm10);
    const R eqmc11 = evlm11 - s13 * (evlm11 - eqm11);
    const R eqmc12 = evlm12 - s13 * (evlm12 - eqm12);
    const R eqmc13 = evlm13 - s16 * (evlm13 - eqm13);
    const R eqmc14 = evlm14 - s16 * (evlm14 - eqm14);
    const R eqmc15 = evlm15 - s16 * (evlm15 - eqm15);

    const R tl1 = val1i * rho9;
    const R tl2 = coef2 * val2i * eqmc1;
    const R tl3 = coef3 * val2i * eqmc1;
    const R tl4 = coef4 * val3i * eqmc2;
    const R tl5 = val3i * eqmc2;
    const R tl6 = val4i * ux9;
    const R tl7 = val5i * eqmc3;
    const R tl8 = val4i * uy9;
    const R tl9 = val5i * eqmc4;
    const R tl10 = val4i * uz9;
    const R tl11 = val5i * eqmc5;
    const R tl12 = val6i * eqmc6;
    const R tl13 = val7i * eqmc7;
    const R tl14 = val8i * eqmc8;
    const R tl15 = val9i * eqmc9;
    const R tl16 = -(coef4i * eqmc10);
    const R tl17 = -(coef4i * eqmc11);
    const R tl18 = -(coef4i * eqmc12);
    const R tl19 = coef3i * eqmc13;
    const R tl20 = coef3i * eqmc14;
    const R tl21 = coef3i * eqmc15;

    f9[0] = tl1 - R(30.0) * val2i * eqmc1 + val8 * val3i * eqmc2;

    const R suma = tl1 + tl2 + tl4;
    const R sumb = tl1 + tl3 + tl5;
    const R sumc = tl6 + coef4 * tl7;
    const R sumd = coef5 * tl12 + coef4 * tl13;
    const R sume = tl8 + coef4 * tl9;
    const R sumf = -tl12 + coef5 * tl13 + tl14 - coef5 * tl15;
    const R sumg = tl10 + coef4 * tl11;
    const R sumh = -tl12 + coef5 * tl13 - tl14 + coef5 * tl15;
    const R sumi = tl12 + tl13 + tl14 + tl15;
    const R sumk = tl12 + tl13 - tl14 - tl15;
    const R sump = -(coef5 * tl12) - coef5 * tl13;
    const R sum67 = tl6 + tl7;
    const R sum89 = tl8 + tl9;
    const R sum1011 = tl10 + tl11;

    f9[1] = suma + sumc + sumd;
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

#pragma unroll
    for (int ip = 0; ip < NPOP; ++ip) f[ip] = (f9[ip] + R(0.5) * Fbar[ip]).v;   // :230,:240
}

// ---- fast collision ------------------------------------------------------------------------
// OWN_MOMENTS = true : main loop -- conserved moments are those of f (rho - rho_shift, j + F/2),
//                      the array arguments are ignored;
// OWN_MOMENTS = false: conserved moments are imposed from (rho_c, ux, uy, uz), exactly like the
//                      reference when its rho/u arrays differ from the moments of f
//                      (pre-relaxation, collision.f90:157,162,164,166).
template <bool OWN_MOMENTS>
D3Q_HD void collide_fast(double (&f)[NPOP], double rho_c, double ux, double uy, double uz,
                         double Fx, double Fy, double Fz, double rho_shift, const Mrt &c) {
    // opposite-pair sums and differences
    const double sx = f[1] + f[2], dx = f[1] - f[2];
    const double sy = f[3] + f[4], dy = f[3] - f[4];
    const double sz = f[5] + f[6], dz = f[5] - f[6];
    const double sa = f[7] + f[10], da = f[7] - f[10];     // (+,+,0) / (-,-,0)
    const double sb = f[9] + f[8], db = f[9] - f[8];       // (+,-,0) / (-,+,0)
    const double sc = f[11] + f[14], dc = f[11] - f[14];   // (+,0,+) / (-,0,-)
    const double sd = f[13] + f[12], dd = f[13] - f[12];   // (+,0,-) / (-,0,+)
    const double se = f[15] + f[18], de = f[15] - f[18];   // (0,+,+) / (0,-,-)
    const double sg = f[17] + f[16], dg = f[17] - f[16];   // (0,+,-) / (0,-,+)

    // odd moments
    const double A = da + db, B = da - db, Cx = dc + dd, D = dc - dd, E = de + dg, G = de - dg;
    const double Tx = A + Cx, Ty = B + E, Tz = D + G;
    const double jx = dx + Tx, jy = dy + Ty, jz = dz + Tz;
    const double mx = A - Cx, my = E - B, mz = D - G;

    // even moments
    const double Sxy = sa + sb, Sxz = sc + sd, Syz = se + sg;
    const double syz = sy + sz, S1 = sx + syz, Sxx = Sxy + Sxz, S2 = Sxx + Syz;
    const double rho = f[0] + S1 + S2;
    const double e = -30.0 * f[0] - 11.0 * S1 + 8.0 * S2;
    const double ep = 12.0 * f[0] - 4.0 * S1 + S2;
    const double tt = Sxx - 2.0 * Syz;
    const double pxx = 2.0 * sx - syz + tt;
    const double pixx = -4.0 * sx + 2.0 * syz + tt;
    const double dyz = sy - sz, dxx = Sxy - Sxz;
    const double pww = dyz + dxx, piww = dxx - 2.0 * dyz;
    const double pxy = sa - sb, pyz = se - sg, pxz = sc - sd;

    // conserved moments seen by the equilibria, and their imposed change
    double r_, jX, jY, jZ;          // delta'_rho, delta'_j (already divided by 19 and 10)
    double jtx, jty, jtz;           // u - F/2
    if (OWN_MOMENTS) {
        rho_c = rho - rho_shift;
        ux = jx + 0.5 * Fx; uy = jy + 0.5 * Fy; uz = jz + 0.5 * Fz;
        r_ = -rho_shift * (1.0 / 19.0);
        jX = Fx * 0.1; jY = Fy * 0.1; jZ = Fz * 0.1;
        jtx = jx; jty = jy; jtz = jz;
    } else {
        r_ = (rho_c - rho) * (1.0 / 19.0);
        jtx = ux - 0.5 * Fx; jty = uy - 0.5 * Fy; jtz = uz - 0.5 * Fz;
        jX = (jtx + Fx - jx) * 0.1; jY = (jty + Fy - jy) * 0.1; jZ = (jtz + Fz - jz) * 0.1;
    }
    const double uxx = ux * ux, uyy = uy * uy, uzz = uz * uz;
    const double u2 = uxx + uyy + uzz;
    const double uF = ux * Fx + uy * Fy + uz * Fz;

    // relaxation: delta'_k = (mF_k - s_k (m_k + mF_k/2 - meq_k)) / N_k
    const double ne = e + 11.0 * rho_c - 19.0 * (u2 - uF);
    const double e_ = (38.0 * uF - c.s1 * ne) * (1.0 / 2394.0);
    const double np = ep - 5.5 * uF - c.omegepsl * rho_c - c.omegepslj * u2;
    const double p_ = (-11.0 * uF - c.s2 * np) * (1.0 / 252.0);

    const double nqx = (-4.0 * dx + Tx) + (2.0 / 3.0) * jtx;
    const double nqy = (-4.0 * dy + Ty) + (2.0 / 3.0) * jty;
    const double nqz = (-4.0 * dz + Tz) + (2.0 / 3.0) * jtz;
    const double qX = ((-2.0 / 3.0) * Fx - c.s4 * nqx) * (1.0 / 40.0);
    const double qY = ((-2.0 / 3.0) * Fy - c.s4 * nqy) * (1.0 / 40.0);
    const double qZ = ((-2.0 / 3.0) * Fz - c.s4 * nqz) * (1.0 / 40.0);

    const double gx = ux * Fx, gy = uy * Fy, gz = uz * Fz;
    const double gxx = 2.0 * gx - gy - gz;            // mF(3pxx)/2
    const double gww = gy - gz;                       // mF(pww)/2
    const double eqxx = 2.0 * uxx - uyy - uzz;
    const double eqww = uyy - uzz;
    const double na = pxx + gxx - eqxx;
    const double a_ = (2.0 * gxx - c.s9 * na) * (1.0 / 36.0);
    const double nb = pixx - 0.5 * gxx - c.omegxx * eqxx;
    const double b_ = (-gxx - c.s10 * nb) * (1.0 / 72.0);
    const double nc = pww + gww - eqww;
    const double c_ = (2.0 * gww - c.s9 * nc) * (1.0 / 12.0);
    const double nd = piww - 0.5 * gww - c.omegxx * eqww;
    const double d_ = (-gww - c.s10 * nd) * (1.0 / 24.0);

    const double gxy = ux * Fy + uy * Fx, gyz = uy * Fz + uz * Fy, gxz = ux * Fz + uz * Fx;
    const double xy_ = (gxy - c.s13 * (pxy + 0.5 * gxy - ux * uy)) * 0.25;
    const double yz_ = (gyz - c.s13 * (pyz + 0.5 * gyz - uy * uz)) * 0.25;
    const double xz_ = (gxz - c.s13 * (pxz + 0.5 * gxz - ux * uz)) * 0.25;

    const double k16 = -c.s16 * 0.125;
    const double mX = k16 * mx, mY = k16 * my, mZ = k16 * mz;

    // back-transform: f_i += sum_k M_ki delta'_k
    f[0] += r_ - 30.0 * e_ + 12.0 * p_;

    const double base_ax = r_ - 11.0 * e_ - 4.0 * p_;
    const double ab2 = 2.0 * b_ - a_;                  // -a + 2b
    const double cd2 = c_ - 2.0 * d_;
    const double evx = base_ax - 2.0 * ab2;            // 2a - 4b
    const double evy = base_ax + ab2 + cd2;
    const double evz = base_ax + ab2 - cd2;
    const double odx = jX - 4.0 * qX, ody = jY - 4.0 * qY, odz = jZ - 4.0 * qZ;
    f[1] += evx + odx; f[2] += evx - odx;
    f[3] += evy + ody; f[4] += evy - ody;
    f[5] += evz + odz; f[6] += evz - odz;

    const double base_dg = r_ + 8.0 * e_ + p_;
    const double apb = a_ + b_, cpd = c_ + d_;
    const double evxy = base_dg + apb + cpd;
    const double evxz = base_dg + apb - cpd;
    const double evyz = base_dg - 2.0 * apb;
    const double X = jX + qX, Y = jY + qY, Z = jZ + qZ;

    const double o7 = X + Y + (mX - mY), o9 = X - Y + (mX + mY);
    const double o11 = X + Z + (mZ - mX), o13 = X - Z - (mX + mZ);
    const double o15 = Y + Z + (mY - mZ), o17 = Y - Z + (mY + mZ);

    const double sp7 = evxy + xy_, sp9 = evxy - xy_;
    f[7] += sp7 + o7;   f[10] += sp7 - o7;
    f[9] += sp9 + o9;   f[8] += sp9 - o9;
    const double sp11 = evxz + xz_, sp13 = evxz - xz_;
    f[11] += sp11 + o11; f[14] += sp11 - o11;
    f[13] += sp13 + o13; f[12] += sp13 - o13;
    const double sp15 = evyz + yz_, sp17 = evyz - yz_;
    f[15] += sp15 + o15; f[18] += sp15 - o15;
    f[17] += sp17 + o17; f[16] += sp17 - o17;
}

// rho = sum f in index order, rhoupdat (collision.f90:475-478)
D3Q_HD double rho_index_order(const double (&f)[NPOP]) {
    R r(f[0]);
#pragma unroll
    for (int ip = 1; ip < NPOP; ++ip) r = r + R(f[ip]);
    return r.v;
}

}  // namespace d3q
