"""Node-local algebra of the CUDA kernels (csrc/collide.cuh), compiled for the host with g++.

strict arithmetic must be BIT-IDENTICAL to the oracle (same expressions, no contraction);
fast arithmetic (the production kernels) must agree to rounding.  Streaming is a pure copy, so
host-collided populations pushed through oracle/textbook.stream must equal the oracle's f.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import textbook as tb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "d3q19-single-phase_b200", "csrc")


@pytest.fixture(scope="module")
def host():
    out = os.path.join(ROOT, "tests", "host", "libcollide_host.so")
    src = os.path.join(ROOT, "tests", "host", "collide_host.cpp")
    deps = [src, os.path.join(CSRC, "collide.cuh"), os.path.join(CSRC, "lattice.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", CSRC,
                        "-o", out, src], check=True)
    L = C.CDLL(out)
    dp = C.POINTER(C.c_double)
    L.collide_nodes.argtypes = [C.c_int, C.c_long, dp, dp, dp, dp, dp] + [C.c_double] * 4 + [dp]
    L.moments_nodes.argtypes = [C.c_long, dp] + [C.c_double] * 3 + [dp] * 4
    return L


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _mrt(p):
    return np.array([p.s1, p.s2, p.s4, p.s9, p.s10, p.s13, p.s16, p.omegepsl, p.omegepslj, p.omegxx])


def _collide(host, mode, f, p, F, macro=None, shift=0.0):
    g = np.ascontiguousarray(f.reshape(-1, 19).copy())
    n = g.shape[0]
    z = np.zeros(n)
    rho, ux, uy, uz = [np.ascontiguousarray(a.reshape(-1)) for a in macro] if macro else (z, z, z, z)
    host.collide_nodes(mode, n, _p(g), _p(rho), _p(ux), _p(uy), _p(uz), F[0], F[1], F[2], shift, _p(_mrt(p)))
    return g.reshape(f.shape)


@pytest.mark.parametrize("laminar,mrt", [(False, 1), (True, 2), (False, 3)])
def test_strict_is_bit_identical_to_oracle(oracle, host, laminar, mrt):
    nx, ny, nz = 11, 5, 6
    w, p = oracle.make_initial_state(nx, ny, nz, laminar=laminar, noise=True, MRTtype=mrt)
    rng = np.random.default_rng(11)
    f0 = w.get_f() + 1e-3 * rng.normal(size=(nz, ny, nx, 19))
    w.set_f(f0)
    F = (0.0, p.force_in_y * p.force_mag, 0.0)
    for _ in range(3):
        w.macrovar()
        w.collision_MRT()
        f0 = tb.stream(_collide(host, 0, f0, p, F))
        assert np.array_equal(f0, w.get_f())


def test_strict_prerelax_is_bit_identical(oracle, host):
    nx, ny, nz = 8, 4, 4
    w, p = oracle.make_initial_state(nx, ny, nz, laminar=False, noise=True)
    F = (0.0, p.force_in_y * p.force_mag, 0.0)
    f = w.get_f()
    u = [w.get(k) for k in ("ux", "uy", "uz")]
    for _ in range(3):
        w.rhoupdat()
        rho = w.get("rho")
        w.collision_MRT()
        f = tb.stream(_collide(host, 1, f, p, F, macro=(rho, *u)))
        assert np.array_equal(f, w.get_f())


@pytest.mark.parametrize("mrt", [1, 2, 3])
@pytest.mark.parametrize("F", [(0.0, 3e-6, 0.0), (1e-5, -2e-5, 3e-5)])
def test_fast_matches_strict_to_rounding(oracle, host, mrt, F):
    p = oracle.make_para(64, 32, 32, laminar=False, MRTtype=mrt)
    rng = np.random.default_rng(5)
    n = 4096
    u = 0.05 * rng.normal(size=(3, n))
    f = np.empty((n, 19))
    usq = 1.5 * (u**2).sum(0)
    for i in range(19):
        G = tb.CX[i] * u[0] + tb.CY[i] * u[1] + tb.CZ[i] * u[2]
        f[:, i] = tb.W[i] * (3 * G + 4.5 * G * G - usq)
    f += 1e-3 * rng.normal(size=f.shape)
    scale = np.max(np.abs(f))
    a = _collide(host, 0, f, p, F)
    b = _collide(host, 2, f, p, F)
    assert np.max(np.abs(a - b)) < 2e-15 * scale
    # rho shift of avedensity (collision.f90:505-511)
    a = _collide(host, 0, f, p, F, shift=1.25e-4)
    b = _collide(host, 2, f, p, F, shift=1.25e-4)
    assert np.max(np.abs(a - b)) < 2e-15 * scale
    # imposed conserved moments (pre-relaxation / external arrays)
    macro = (1e-3 * rng.normal(size=n), *(0.05 * rng.normal(size=(3, n))))
    a = _collide(host, 1, f, p, F, macro=macro)
    b = _collide(host, 3, f, p, F, macro=macro)
    assert np.max(np.abs(a - b)) < 4e-15 * max(scale, 0.05)
    # and against the independent matrix form
    t = tb.collide(p, f, *macro, np.full(n, F[0]), np.full(n, F[1]), np.full(n, F[2]))
    assert np.max(np.abs(b - t)) < 1e-14 * max(scale, 0.05)


def test_moments_strict_match_macrovar(oracle, host):
    nx, ny, nz = 7, 3, 4
    w, p = oracle.make_initial_state(nx, ny, nz, laminar=False, noise=True)
    f = w.get_f()
    w.macrovar()
    n = nx * ny * nz
    out = [np.empty(n) for _ in range(4)]
    host.moments_nodes(n, _p(np.ascontiguousarray(f.reshape(-1, 19))), 0.0, p.force_in_y * p.force_mag, 0.0,
                       *[_p(o) for o in out])
    for o, k in zip(out, ("rho", "ux", "uy", "uz")):
        assert np.array_equal(o.reshape(nz, ny, nx), w.get(k))
