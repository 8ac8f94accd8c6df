"""BASELINE.json's full size (configs[1]: 512x256x256, 33.5 M nodes, turbulent set) on the GPU: three steps against the
REFERENCE ITSELF (the translated Fortran, ~1 s per step on the host cores, test_reference_bit_for_bit_at_full_size),
and, for longer runs, size-independent properties:

* the two storage schemes are independent implementations of the same map (one-step pull into a second array
  vs. in-place even/odd): in STRICT arithmetic they must stay BIT-IDENTICAL, seen through probes at corner / wall /
  bulk nodes and through the plane sums of every x-plane (a sum over all 33.5 M nodes: any mis-indexed node shows);
* download -> upload into the other scheme is the identity (gather / scatter / un-stream kernels at full size);
* mass is conserved by collision, forcing, streaming and the bounce-back walls;
* the production (FAST, FMA) arithmetic stays within BASELINE.json's 1e-12 (1 step) / 1e-9 (here 100 steps) of STRICT.
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


# ---- the reference itself at full size ------------------------------------------------------------------------------
def _reference_run(nx, ny, nz, steps_at):
    """the translated reference (oracle/_ref/libref.so: collision.f90 / para.f90 / initial.f90 machine-translated to C,
    IEEE evaluation of the source order; the hand restatement is bit-identical to it, tests/test_oracle_ref.py) on all
    host cores, one thread per emulated MPI rank: f0 and f after each step count in `steps_at`"""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libref.so was not built (it is built where /root/reference is mounted and travels with the repo)")
    ncores = os.cpu_count() or 1
    npz = 1
    while npz * 2 <= ncores and nz % (npz * 2) == 0 and (npz * 2) ** 2 <= ncores * 2:
        npz *= 2
    npy = max(1, ncores // npz)
    while ny % npy:
        npy -= 1
    w = ref.RefWorld(nx, ny, nz, nprocY=npy, nprocZ=npz, laminar=False, a9=0.3, **({} if nx >= 512 else dict(ustar=0.0025)))
    w.run("initvel")                                   # main.f90:58 (log-law + the A9 perturbation block)
    rng = np.random.default_rng(54321)
    for k in ("ux", "uy", "uz"):                       # seeded noise so that every population is distinct
        w.set(k, w.get(k) + 1e-3 * w.scalar("ustar") * (2.0 * rng.random((nz, ny, nx)) - 1.0))
    w.run("forcing"); w.run("initpop"); w.run("macrovar")          # main.f90:61,65,136
    yield w.get_f()
    done = 0
    for n in steps_at:
        w.loop("collision_mrt", "macrovar", n - done)              # main.f90:157-161
        done = n
        yield w.get_f()
    w.close()


@pytest.mark.parametrize("shape,schemes", [((NX, NY, NZ), ("aa", "ab")), ((2 * NX, 4 * NY, max(NZ // 16, 4)), ("aa",))],
                         ids=["configs1", "c4_plane_1024x1024"])
def test_reference_bit_for_bit_at_full_size(shape, schemes):
    """configs[1] (512x256x256) and a 1024x1024-wide slab of configs[3]'s planes (in place, as C4 must run): 3 steps of the
    reference's own code and of the GPU from the same bits -- STRICT arithmetic bit for bit, production arithmetic
    within BASELINE.json's 1e-12 after one step"""
    nx, ny, nz = shape
    kw = {} if nx >= 512 else dict(ustar=0.0025, ystar=0.0036 / 0.0025, force_in_y=2.0 * 0.0025 * 0.0025 / nx)
    gen = _reference_run(nx, ny, nz, (1, 3))
    f0 = next(gen)
    sims = []
    for name in schemes:
        for mm in (capi.MATH_STRICT, capi.MATH_FAST):
            if mm == capi.MATH_FAST and name != schemes[-1]:
                continue                                           # one production-arithmetic run is enough
            sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, scheme=capi.SCHEME_AA if name == "aa" else capi.SCHEME_AB,
                                  math_mode=mm, allocate_host=False, **kw)
            sim.FORCING()
            sim.upload_f(f0)
            sims.append((name, mm, sim))
    del f0
    out = np.empty((nz, ny, nx, 19))
    done = 0
    for n in (1, 3):
        ref_f = next(gen)
        scale = float(np.max(np.abs(ref_f)))
        for name, mm, sim in sims:
            sim.run_device(n - done)
            sim.download_f(out)
            if mm == capi.MATH_STRICT:
                assert np.array_equal(out, ref_f), (name, n)
            else:
                np.subtract(out, ref_f, out=out)
                err = float(np.max(np.abs(out))) / scale
                assert err < (1e-12 if n == 1 else 1e-11), (name, n, err)
        done = n
        del ref_f
    for _, _, sim in sims:
        sim.close()
    for _ in gen:
        pass
