"""BASELINE.json's full size (configs[1]: 512x256x256, 33.5 M nodes, turbulent set) on the GPU, through
size-independent properties -- the oracle needs ~0.5 s per step and 20 GB of host arrays at this size, so the
direct comparisons stay at the sizes of test_gpu_parity.py / test_golden.py:

* the two storage schemes are independent implementations of the same map (one-step pull into a second array
  vs. in-place even/odd): in STRICT arithmetic they must stay BIT-IDENTICAL, seen through probes at corner / wall /
  bulk nodes and through the plane sums of every x-plane (a sum over all 33.5 M nodes: any mis-indexed node shows);
* download -> upload into the other scheme is the identity (gather / scatter / un-stream kernels at full size);
* mass is conserved by collision, forcing, streaming and the bounce-back walls;
* the production (FAST, FMA) arithmetic stays within BASELINE.json's 1e-12 (1 step) / 1e-9 (here 100 steps) of STRICT.

The file sorts last on purpose: it was written in a session without GPU time.
"""
import numpy as np
import pytest

import __graft_entry__ as entry

pytestmark = pytest.mark.gpu

pkg = entry.load_package()
capi = pkg.capi

import os

# D3Q19_TEST_FULLSIZE=AxBxC shrinks the case (tests/test_hostsim_capi.py runs this file's logic on the host-sim build)
NX, NY, NZ = (int(t) for t in os.environ.get("D3Q19_TEST_FULLSIZE", "512x256x256").split("x"))
PROBES = [(1, 1, 1), (NX, NY, NZ), (1, NY, 1), (NX, 1, NZ), (NX // 2, NY // 2, NZ // 2), (2, min(17, NY), NZ - 1), (NX - 1, NY, 2),
          (NX // 4, 1, NZ // 2), (NX // 4 + 1, NY // 2, NZ)]
EXACT_ROWS = [0, 1, 2, 9, 11]          # sums of ux, uy, uz, rho (pure additions of strict moments) and the node count


# a shrunk channel keeps the wall units of the 512-wide one (para.f90:64-66 would give Ma ~ 1 at nx = 64)
SHRUNK = {} if NX >= 512 else dict(ustar=0.0025, ystar=0.0036 / 0.0025, force_in_y=2.0 * 0.0025 * 0.0025 / NX)


def start(scheme, math_mode):
    sim = pkg.ChannelFlow(NX, NY, NZ, laminar=False, scheme=scheme, math_mode=math_mode, allocate_host=False, **SHRUNK)
    sim.FORCING()
    sim.init_channel_device(A9=0.3, noise_amp=1e-3 * sim.v.ustar, seed=54321)
    return sim


def fingerprint(sim):
    return sim.profiles2(), np.array([sim.probe(*p) for p in PROBES])


def same_bits(a, b):
    pa, qa = a
    pb, qb = b
    return bool(np.array_equal(qa, qb) and np.array_equal(pa[EXACT_ROWS], pb[EXACT_ROWS]))


def row_scales(p):
    """what a plane sum is measured against: sums that cancel (ux, uz, rho, the cross products) are pure rounding
    noise, so linear rows are scaled by the largest |sum uy| and quadratic rows by the largest sum uy^2"""
    lin, quad = np.max(np.abs(p[1])), np.max(np.abs(p[4]))
    return np.array([lin, lin, lin, quad, quad, quad, quad, quad, quad, lin, quad])[:, None]


def close_products(a, b, tol=1e-13):
    pa, pb = a[0][:11], b[0][:11]
    return bool(np.max(np.abs(pa - pb) / row_scales(pb)) < tol)


def test_schemes_stay_bit_identical_at_full_size():
    # both runs start from the SAME bits: the device initialisation is not strict arithmetic, so the in-place run's
    # field goes to the two-array run through the host (which is also the gather / scatter / un-stream path)
    aa = start(capi.SCHEME_AA, capi.MATH_STRICT)
    ab = pkg.ChannelFlow(NX, NY, NZ, laminar=False, scheme=capi.SCHEME_AB, math_mode=capi.MATH_STRICT, allocate_host=False,
                         **SHRUNK)
    ab.FORCING()
    f = np.empty((NZ, NY, NX, 19))
    aa.download_f(f)
    ab.upload_f(f)
    del f
    fa, fb = fingerprint(aa), fingerprint(ab)
    assert same_bits(fa, fb) and close_products(fa, fb)
    assert fa[0][11].sum() == NX * NY * NZ
    assert np.ptp(fa[1][:, 2]) > 0                       # the probes see different velocities: not a trivial field
    for n in (1, 2, 7):                                    # odd and even counts: both in-place phases are read
        aa.run_device(n); ab.run_device(n)
        fa, fb = fingerprint(aa), fingerprint(ab)
        assert same_bits(fa, fb), n
        assert close_products(fa, fb), n
    aa.close(); ab.close()


def test_download_upload_roundtrip_at_full_size():
    aa = start(capi.SCHEME_AA, capi.MATH_STRICT)
    aa.run_device(3)                                       # in-place storage in its swapped phase
    f = np.empty((NZ, NY, NX, 19))
    aa.download_f(f)
    ab = pkg.ChannelFlow(NX, NY, NZ, laminar=False, scheme=capi.SCHEME_AB, math_mode=capi.MATH_STRICT, allocate_host=False,
                         **SHRUNK)
    ab.FORCING()
    ab.upload_f(f)
    assert same_bits(fingerprint(aa), fingerprint(ab))
    g = np.empty_like(f)
    ab.download_f(g)
    assert np.array_equal(f, g)
    del g
    aa.run_device(2); ab.run_device(2)
    assert same_bits(fingerprint(aa), fingerprint(ab))
    aa.close(); ab.close()


@pytest.mark.parametrize("scheme", [capi.SCHEME_AA, capi.SCHEME_AB])
def test_mass_is_conserved_at_full_size(scheme):
    sim = start(scheme, capi.MATH_FAST)
    m0 = sim.profiles2()[9].sum()
    sim.run_device(100)
    p = sim.profiles2()
    assert np.all(np.isfinite(p))
    # one population of one node is ~1e-3: a lost or doubled value anywhere would show 5 orders above this bound
    assert abs(p[9].sum() - m0) < 1e-8, (p[9].sum(), m0)
    sim.close()


@pytest.mark.parametrize("scheme", [capi.SCHEME_AA, capi.SCHEME_AB])
def test_fast_arithmetic_tracks_strict_at_full_size(scheme):
    strict, fast = start(scheme, capi.MATH_STRICT), start(scheme, capi.MATH_FAST)
    strict.run_device(1); fast.run_device(1)
    (ps, qs), (pf, qf) = fingerprint(strict), fingerprint(fast)
    scale = np.max(np.abs(qs))
    assert np.max(np.abs(qs - qf)) < 1e-12 * scale         # BASELINE.json: 1 step
    strict.run_device(99); fast.run_device(99)
    (ps, qs), (pf, qf) = fingerprint(strict), fingerprint(fast)
    assert np.max(np.abs(qs - qf)) < 1e-9 * np.max(np.abs(qs))
    rel = np.max(np.abs(ps[:11] - pf[:11]) / row_scales(ps), axis=1)
    assert np.max(rel) < 1e-9, rel
    strict.close(); fast.close()
