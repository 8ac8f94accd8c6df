"""Independent "textbook" restatement of one Channel-Flow time step (numpy, matrix form).

TEST INFRASTRUCTURE ONLY (see oracle/d3q19_oracle.h).  This is deliberately NOT a
transcription of collision.f90: it builds the d'Humieres et al. (2002) D3Q19 moment matrix
from its polynomial definition, relaxes in moment space, and streams with array rolls plus
half-way bounce-back.  It exists to cross-check the line-by-line C oracle
(SURVEY.md fact 5 / fact 6): the two must agree to rounding (~1e-16 relative).

Arrays are numpy C-order f[iz, iy, ix, ip], a[iz, iy, ix]; lattice ordering is the
reference's (para.f90:178-203).
"""
import numpy as np

CX = np.array([0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0])
CY = np.array([0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1])
CZ = np.array([0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1])
OPP = np.array([0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15])
W = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12)


def moment_matrix():
    """Rows in d'Humieres (2002) order: rho, e, eps, jx, qx, jy, qy, jz, qz, 3pxx, 3pixx,
    pww, piww, pxy, pyz, pxz, mx, my, mz."""
    c2 = (CX**2 + CY**2 + CZ**2).astype(float)
    cx, cy, cz = CX.astype(float), CY.astype(float), CZ.astype(float)
    rows = [
        np.ones(19),
        19 * c2 - 30,
        (21 * c2**2 - 53 * c2 + 24) / 2,
        cx, (5 * c2 - 9) * cx,
        cy, (5 * c2 - 9) * cy,
        cz, (5 * c2 - 9) * cz,
        3 * cx**2 - c2, (3 * c2 - 5) * (3 * cx**2 - c2),
        cy**2 - cz**2, (3 * c2 - 5) * (cy**2 - cz**2),
        cx * cy, cy * cz, cx * cz,
        (cy**2 - cz**2) * cx, (cz**2 - cx**2) * cy, (cx**2 - cy**2) * cz,
    ]
    return np.array(rows)


def relaxation_rates(p):
    """Diagonal of S in the row order above, from the reference's named rates
    (collision.f90:140-154: s1,s2,s4,s4,s4,s9,s10,s9,s10,s13 x3,s16 x3)."""
    return np.array([0, p.s1, p.s2, 0, p.s4, 0, p.s4, 0, p.s4, p.s9, p.s10, p.s9, p.s10,
                     p.s13, p.s13, p.s13, p.s16, p.s16, p.s16])


def equilibrium_moments(p, rho, ux, uy, uz):
    """Incompressible (delta-rho) equilibria, collision.f90:86-101, as a (19, ...) array."""
    u2 = ux * ux + uy * uy + uz * uz
    pxx3 = 2 * ux * ux - uy * uy - uz * uz
    pww = uy * uy - uz * uz
    z = np.zeros_like(rho)
    return np.array([
        rho, -11 * rho + 19 * u2, p.omegepsl * rho + p.omegepslj * u2,
        ux, -2 / 3 * ux, uy, -2 / 3 * uy, uz, -2 / 3 * uz,
        pxx3, p.omegxx * pxx3, pww, p.omegxx * pww,
        ux * uy, uy * uz, ux * uz, z, z, z])


def force_populations(ux, uy, uz, fx, fy, fz):
    """Fbar_i = w_i (3 c.F + 9 (c.F)(c.u) - 3 u.F), with Fbar_0 = -u.F (collision.f90:68-82)."""
    uF = ux * fx + uy * fy + uz * fz
    out = np.empty(ux.shape + (19,))
    for i in range(19):
        cF = CX[i] * fx + CY[i] * fy + CZ[i] * fz
        cu = CX[i] * ux + CY[i] * uy + CZ[i] * uz
        out[..., i] = W[i] * (3 * cF + 9 * cF * cu - 3 * uF)
    return out


def moments(f, fx, fy, fz):
    """macrovar's fluid branch (collision.f90:394-418): rho = sum f, u = sum c f + F/2."""
    rho = f.sum(-1)
    ux = (f * CX).sum(-1) + fx / 2
    uy = (f * CY).sum(-1) + fy / 2
    uz = (f * CZ).sum(-1) + fz / 2
    return rho, ux, uy, uz


def collide(p, f, rho, ux, uy, uz, fx, fy, fz):
    """f* = Minv [ conserved from arrays ; m - S (m - meq) ] + Fbar/2 with m = M (f + Fbar/2)."""
    M = moment_matrix()
    Minv = np.linalg.inv(M)
    S = relaxation_rates(p)
    Fbar = force_populations(ux, uy, uz, fx, fy, fz)
    f9 = f + 0.5 * Fbar
    m = np.einsum("ki,...i->k...", M, f9)
    meq = equilibrium_moments(p, rho, ux, uy, uz)
    mstar = m - S.reshape((19,) + (1,) * rho.ndim) * (m - meq)
    # conserved rows are taken from the arrays, not from f9 (collision.f90:157,162,164,166)
    mstar[0], mstar[3], mstar[5], mstar[7] = rho, ux, uy, uz
    return np.einsum("ik,k...->...i", Minv, mstar) + 0.5 * Fbar


def stream(fstar, solid=None):
    """Pull streaming, periodic in y and z, half-way bounce-back at the x walls
    (canonical form)."""
    nz, ny, nx, _ = fstar.shape
    out = np.empty_like(fstar)
    for i in range(19):
        src = np.roll(fstar[..., i], shift=(CZ[i], CY[i]), axis=(0, 1))
        if CX[i] == 0:
            out[..., i] = src
        elif CX[i] == 1:
            out[:, :, 1:, i] = src[:, :, :-1]
            out[:, :, 0, i] = fstar[:, :, 0, OPP[i]]
        else:
            out[:, :, :-1, i] = src[:, :, 1:]
            out[:, :, -1, i] = fstar[:, :, -1, OPP[i]]
    return out


def step(p, f, fx, fy, fz, macro=None):
    """One collision_MRT call.  macro=None: moments of f (main loop, fact 7);
    otherwise the (rho,ux,uy,uz) arrays the reference would read."""
    if macro is None:
        macro = moments(f, fx, fy, fz)
    return stream(collide(p, f, *macro, fx, fy, fz))


def poiseuille_startup(nx, ustar, visc, istep, nterms=26):
    """The reference's analytic start-up solution, saveload.f90:921-933: returns
    (uut, uuss/ustar) for i = 1..nx/2 (both normalised by ustar)."""
    i = np.arange(1, nx // 2 + 1)
    time1 = istep * visc / ((nx / 2.0) ** 2)
    xx0 = np.abs(i - 0.5 - nx / 2.0)
    uuss = (4 * ustar) / (nx**2) * ((nx / 2.0) ** 2 - xx0**2)
    uut = np.zeros_like(xx0)
    for n in range(nterms):
        uut += (4 * (-1) ** n) / (np.pi * (n + 0.5)) ** 3 * np.exp(-((n + 0.5) ** 2) * np.pi**2 * time1) \
            * np.cos((n + 0.5) * np.pi * (xx0 / (nx / 2.0)))
    uut = (1 - (xx0 / (nx / 2.0)) ** 2) - uut
    return uut, uuss / ustar
