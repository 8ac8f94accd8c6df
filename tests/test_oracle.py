"""Pins the CPU oracle (oracle/d3q19_oracle.c) before anything is compared against it.

The reference ships no tests or golden vectors and cannot be compiled
here, so the oracle is pinned to: (a) the analytic Poiseuille known-answers the reference
itself embeds (saveload.f90:921-935), (b) an independent matrix-form restatement
(oracle/textbook.py), (c) the reference's own invariants: results independent of the
nprocY x nprocZ decomposition, mass/momentum identities of the forcing,
LBGK recovery for MRTtype=2 (para.f90:121-130 "To recover LBGK").
"""
import numpy as np
import pytest

from oracle import textbook as tb


def _relerr(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def test_para_constants(oracle):
    # para.f90:75-88 laminar set with nx=64; para.f90:61-70 turbulent set with nx=512
    p = oracle.make_para(64, 32, 32, laminar=True)
    assert p.MRTtype == 2 and p.ivel == 0
    assert p.visc == 2.0 * 0.05 * 64 / 20
    assert p.force_in_y == 8.0 * p.visc * 0.05 / 64.0**2
    assert p.s1 == p.s9 == 1.0 / (3.0 * p.visc + 0.5)
    assert (p.omegepsl, p.omegepslj, p.omegxx) == (3.0, -5.5, -0.5)
    t = oracle.make_para(512, 256, 256, laminar=False)
    assert t.MRTtype == 1 and t.ivel == 1 and t.visc == 0.0036
    assert t.ustar == 2.0 * 180.0 * 0.0036 / 512
    assert (t.s1, t.s2, t.s4, t.s10, t.s16) == (1.5, 1.4, 1.2, 1.4, 1.98)
    assert t.omegepslj == -475.0 / 63.0
    assert list(t.ipopp) == [0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15]
    for i in range(19):
        j = t.ipopp[i]
        assert (t.cix[i], t.ciy[i], t.ciz[i]) == (-t.cix[j], -t.ciy[j], -t.ciz[j])
    assert sorted(list(t.ipswap) + list(t.ipstay)) == list(range(19))


def test_moment_matrix_is_orthogonal_with_reference_norms():
    # the reference divides by 19, 2394, 252, 10, 40, 36, 72, 12, 24, 4, 8 (para.f90:152-160)
    M = tb.moment_matrix()
    G = M @ M.T
    assert np.allclose(G, np.diag(np.diag(G)))
    assert list(np.diag(G)) == [19, 2394, 252, 10, 40, 10, 40, 10, 40, 36, 72, 12, 24, 4, 4, 4, 8, 8, 8]


def test_forcing_moments():
    # sum Fbar = 0, sum c Fbar = F  (SURVEY.md section 4 conservation identities)
    rng = np.random.default_rng(0)
    u = rng.normal(size=(3, 4)) * 0.05
    F = rng.normal(size=(3, 4)) * 1e-3
    Fb = tb.force_populations(*u, *F)
    assert np.allclose(Fb.sum(-1), 0, atol=1e-18)
    assert np.allclose((Fb * tb.CX).sum(-1), F[0], atol=1e-18)
    assert np.allclose((Fb * tb.CY).sum(-1), F[1], atol=1e-18)
    assert np.allclose((Fb * tb.CZ).sum(-1), F[2], atol=1e-18)


@pytest.mark.parametrize("mrt", [1, 2, 3])
def test_oracle_matches_textbook_one_step(oracle, mrt):
    nx, ny, nz = 12, 6, 5
    w, p = oracle.make_initial_state(nx, ny, nz, laminar=False, noise=True, MRTtype=mrt)
    # make every population distinct and non-equilibrium
    rng = np.random.default_rng(7)
    f0 = w.get_f() + 1e-3 * rng.normal(size=(nz, ny, nx, 19))
    w.set_f(f0)
    w.macrovar()
    fx, fy, fz = w.get("fx"), w.get("fy"), w.get("fz")
    rho, ux = w.get("rho"), w.get("ux")
    trho, tux, _, _ = tb.moments(f0, fx, fy, fz)
    scale = np.max(np.abs(f0))          # rho, u are cancelling sums of O(scale) terms
    assert np.max(np.abs(rho - trho)) < 1e-14 * scale and np.max(np.abs(ux - tux)) < 1e-14 * scale
    w.collision_MRT()
    f1 = w.get_f()
    t1 = tb.step(p, f0, fx, fy, fz)
    assert _relerr(f1, t1) < 5e-15
    # mass is conserved exactly by collide+stream with bounce-back
    assert abs(f1.sum() - f0.sum()) < 1e-12


def test_oracle_external_macro_mode_matches_textbook(oracle):
    # pre-relaxation: u frozen at initvel values, rho refreshed by rhoupdat (main.f90:70-90)
    nx, ny, nz = 10, 4, 6
    w, p = oracle.make_initial_state(nx, ny, nz, laminar=False, noise=True)
    ux, uy, uz = w.get("ux"), w.get("uy"), w.get("uz")
    fx, fy, fz = w.get("fx"), w.get("fy"), w.get("fz")
    f = w.get_f()
    for _ in range(3):
        w.rhoupdat()
        rho = w.get("rho")
        assert np.max(np.abs(rho - f.sum(-1))) < 1e-14 * np.max(np.abs(f))
        w.collision_MRT()
        f = tb.step(p, f, fx, fy, fz, macro=(rho, ux, uy, uz))
        assert _relerr(w.get_f(), f) < 1e-13


def test_mrttype2_is_lbgk(oracle):
    # para.f90:121 "To recover LBGK": f* = f9 - (f9 - feq(rho,u))/tau + Fbar/2
    nx, ny, nz = 8, 4, 4
    w, p = oracle.make_initial_state(nx, ny, nz, laminar=True, noise=False)
    rng = np.random.default_rng(3)
    f0 = 1e-3 * rng.normal(size=(nz, ny, nx, 19))
    w.set_f(f0)
    w.macrovar()
    fx, fy, fz = w.get("fx"), w.get("fy"), w.get("fz")
    rho, ux, uy, uz = (w.get(k) for k in ("rho", "ux", "uy", "uz"))
    Fb = tb.force_populations(ux, uy, uz, fx, fy, fz)
    f9 = f0 + 0.5 * Fb
    usq = 1.5 * (ux**2 + uy**2 + uz**2)
    feq = np.empty_like(f0)
    for i in range(19):
        G = tb.CX[i] * ux + tb.CY[i] * uy + tb.CZ[i] * uz
        feq[..., i] = tb.W[i] * (rho + 3 * G + 4.5 * G * G - usq)
    fstar = f9 - (f9 - feq) / p.tau + 0.5 * Fb
    w.collision_MRT()
    assert _relerr(w.get_f(), tb.stream(fstar)) < 1e-13


@pytest.mark.parametrize("grid", [(1, 2), (2, 1), (2, 2), (3, 2), (4, 3)])
def test_decomposition_invariance_is_bitwise(oracle, grid):
    # streaming is a copy and collision is node-local, so the reference's
    # result does not depend on nprocY x nprocZ (including uneven splits, para.f90:233-244).
    nx, ny, nz = 9, 10, 7
    npY, npZ = grid
    w1, _ = oracle.make_initial_state(nx, ny, nz, laminar=False, noise=True)
    wn, _ = oracle.make_initial_state(nx, ny, nz, laminar=False, noise=True, nprocY=npY, nprocZ=npZ)
    # initvel's perturbation offsets assume a uniform split (initial.f90:120,123) but A9 = 0,
    # and the noise is applied on global arrays, so both start identical:
    assert np.array_equal(w1.get_f(), wn.get_f())
    for _ in range(4):
        w1.macrovar(); wn.macrovar()
        w1.collision_MRT(); wn.collision_MRT()
        assert np.array_equal(w1.get_f(), wn.get_f())
    w1.macrovar(); wn.macrovar()
    for k in ("rho", "ux", "uy", "uz"):
        assert np.array_equal(w1.get(k), wn.get(k))


def test_poiseuille_startup_matches_reference_analytic(oracle):
    # The default laminar case of the reference (para.f90:75-88) against the start-up series
    # it prints beside its own profile (saveload.f90:921-935).  Half-way bounce-back places the
    # wall half a spacing outside the first node, for which the scheme is second-order accurate.
    nx, ny, nz = 32, 4, 4
    w, p = oracle.make_initial_state(nx, ny, nz, laminar=True, noise=False)
    assert np.all(w.get_f() == 0.0)        # u = 0, rho = 0 => f = 0 (initial.f90:32-44)
    w.macrovar()
    for istep in range(1, 601):
        w.collision_MRT()
        w.macrovar()
        if istep in (100, 300, 600):
            uy = w.get("uy")[nz // 2, 0, : nx // 2] / p.ustar
            uut, _ = tb.poiseuille_startup(nx, p.ustar, p.visc, istep)
            assert np.max(np.abs(uy - uut)) < 2e-3, istep
    assert np.max(np.abs(w.get("ux"))) < 1e-15 and np.max(np.abs(w.get("uz"))) < 1e-15
    # uniform in y and z
    uyf = w.get("uy")
    assert np.max(np.abs(uyf - uyf[0:1, 0:1, :])) < 1e-16


def test_poiseuille_steady_state(oracle):
    nx = 16
    w, p = oracle.make_initial_state(nx, 2, 2, laminar=True, noise=False)
    w.macrovar()
    for _ in range(4000):
        w.collision_MRT()
        w.macrovar()
    uy = w.get("uy")[0, 0, : nx // 2] / p.ustar
    _, uuss = tb.poiseuille_startup(nx, p.ustar, p.visc, 4000)
    # LBGK + half-way bounce-back: parabolic profile with an O(1/nx^2) constant slip offset
    assert np.max(np.abs(uy - uuss)) < 5e-3
    d2 = np.diff(uy, 2)
    assert np.allclose(d2, -8.0 / nx**2, rtol=2e-5)   # parabolic with the analytic curvature
    # wall shear from the momentum balance tau_w = force_in_y * nx / 2
    assert np.isclose(p.force_in_y * nx / 2, p.visc * 2 * p.ustar * 2 / nx, rtol=1e-12)


def test_avedensity(oracle):
    nx, ny, nz = 6, 4, 4
    w, p = oracle.make_initial_state(nx, ny, nz, laminar=False, noise=True, nprocY=2, nprocZ=2)
    rng = np.random.default_rng(5)
    w.set_f(w.get_f() + 1e-3 * rng.random((nz, ny, nx, 19)))
    w.macrovar()
    rho = w.get("rho")
    mean, nfluid = w.avedensity()
    assert nfluid == nx * ny * nz
    assert np.isclose(mean, rho.mean(), rtol=1e-13)
    assert np.allclose(w.get("rho"), rho - mean, atol=1e-18)
