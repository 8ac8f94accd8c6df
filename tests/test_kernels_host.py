"""The CUDA kernels of csrc/kernels.cuh and csrc/particles.cuh compiled for the HOST (tests/host/kernels_host.cpp + a
small CUDA vocabulary, tests/host/fake/cuda_runtime.h: launches as loop nests, cooperating blocks as fibers) and run
through the launch sequences of d3q19_api.cu: index logic of the three storage phases, wall select, y/z wraps, pitch padding, 32/64-bit indices, the five-population face
pack/unpack, the send-back after an in-place odd step and the stores into a neighbour's array of the peer-memory
and "put" halos -- all bit for bit against the oracle, without a GPU.  The GPU suite (tests/test_gpu_*.py) proves
the same through the C-ABI on the device; this file keeps the kernels' logic checked on the build box.
Test infrastructure only: the product has no host path.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "d3q19-single-phase_b200", "csrc")
HOST = os.path.join(ROOT, "tests", "host")

pkg = entry.load_package()

AA, AB = 0, 1
PACKED, FUSED, FUSED_SPLIT, PUT = 0, 1, 2, 3
MACRO_MAIN, MACRO_PRERELAX, MACRO_EXTERNAL = 0, 1, 2


@pytest.fixture(scope="module")
def hk():
    out = os.path.join(HOST, "libkernels_host.so")
    deps = [os.path.join(HOST, "kernels_host.cpp"), os.path.join(HOST, "fake", "cuda_runtime.h")] + \
           [os.path.join(CSRC, n) for n in ("kernels.cuh", "particles.cuh", "collide.cuh", "lattice.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-I", os.path.join(HOST, "fake"), "-I", CSRC, "-o", out, os.path.join(HOST, "kernels_host.cpp")],
                       check=True)
    L = C.CDLL(out)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
    L.hs_create.restype = vp
    L.hs_create.argtypes = [C.c_int] * 4 + [ip] + [C.c_int] * 4 + [dp] + [C.c_double] * 3 + [C.c_int]
    L.hs_destroy.argtypes = [vp]
    L.hs_upload.argtypes = [vp, dp]
    L.hs_download.argtypes = [vp, dp]
    L.hs_steps.argtypes = [vp, C.c_int, C.c_int]
    L.hs_set_rho_shift.argtypes = [vp, C.c_double]
    L.hs_set_macro.argtypes = [vp, dp, dp, dp, dp]
    L.hs_set_force_field.argtypes = [vp, dp, dp, dp]
    L.hs_set_solid.argtypes = [vp, ip, ip, C.c_int, dp, dp, dp, C.c_double]
    L.hs_macrovar.argtypes = [vp, C.c_int, dp, dp, dp, dp]
    L.hs_vortcalc.argtypes = [vp, dp, dp, dp]
    L.hs_profiles.argtypes = [vp, C.c_int, dp]
    L.hs_init_channel.argtypes = [vp] + [C.c_double] * 4 + [C.c_uint64, C.c_int]
    L.hs_forcingp.argtypes = [vp, C.c_int, C.c_int] + [C.c_double] * 5 + [dp, dp, dp]
    L.hs_avedensity.restype = C.c_double
    L.hs_avedensity.argtypes = [vp, C.POINTER(C.c_longlong)]
    L.hs_prerelax.argtypes = [vp, C.c_double, C.c_int, dp]
    L.hs_diag.argtypes = [vp, C.c_int, dp]
    L.hs_particles_init.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_double, dp, dp, dp, dp, dp]
    L.hs_beads_links.restype = C.c_longlong
    L.hs_beads_links.argtypes = [vp]
    L.hs_particle_step.argtypes = [vp, C.c_int]
    L.hs_get_mask.argtypes = [vp, ip]
    L.hs_get_links.restype = C.c_longlong
    L.hs_get_links.argtypes = [vp, C.c_int, ip, ip, ip, ip, ip, dp]
    L.hs_get_particles.argtypes = [vp, dp, dp, dp, dp, dp]
    return L


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class HostSim:
    """All z-slabs of a run on the host-compiled kernels."""

    def __init__(self, L, p, nranks=1, scheme=AA, strict=True, transport=PACKED, idx64=False, pf_blocks=128, lz=None):
        self.L, self.shape = L, (p.nz, p.ny, p.nx)
        if lz is None:
            lz = [pkg.slab(p.nz, nranks, r)[0] for r in range(nranks)]        # para.f90:241-245
        assert sum(lz) == p.nz
        lz = np.array(lz, dtype=np.int32)
        mrt = np.array([p.s1, p.s2, p.s4, p.s9, p.s10, p.s13, p.s16, p.omegepsl, p.omegepslj, p.omegxx])
        self.h = L.hs_create(p.nx, p.ny, p.nz, nranks, _i(lz), scheme, int(strict), transport, int(idx64), _d(mrt),
                             0.0, p.force_in_y * p.force_mag, 0.0, pf_blocks)          # FORCING, collision.f90:522-524

    def upload(self, f):
        self.L.hs_upload(self.h, _d(np.ascontiguousarray(f)))

    def download(self):
        out = np.full(self.shape + (19,), np.nan)
        self.L.hs_download(self.h, _d(out))
        return out

    def steps(self, n=1, mode=MACRO_MAIN):
        self.L.hs_steps(self.h, n, mode)

    def macrovar(self, rho_only=False):
        o = [np.full(self.shape, np.nan) for _ in range(4)]
        self.L.hs_macrovar(self.h, int(rho_only), *[_d(a) for a in o])
        return o

    def close(self):
        self.L.hs_destroy(self.h)


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def pair(oracle, hk, shape, laminar=False, perturb=0.0, **kw):
    nx, ny, nz = shape
    over = {k: kw.pop(k) for k in ("MRTtype", "ipart") if k in kw}
    w, p = oracle.make_initial_state(nx, ny, nz, laminar=laminar, noise=not laminar, **over)
    if perturb:
        w.set_f(w.get_f() + perturb * np.random.default_rng(99).normal(size=(nz, ny, nx, 19)))
    sim = HostSim(hk, p, **kw)
    sim.upload(w.get_f())
    w.macrovar()
    return w, p, sim


SHAPES = [(11, 5, 6), (16, 1, 1), (130, 3, 2), (23, 4, 3)]


@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("shape", SHAPES)
def test_upload_download_roundtrip(oracle, hk, scheme, shape):
    w, p, sim = pair(oracle, hk, shape, scheme=scheme, perturb=1e-3)
    assert np.array_equal(sim.download(), w.get_f())
    sim.close()


@pytest.mark.parametrize("idx64", [False, True])
@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("shape,laminar,mrt", [((11, 5, 6), False, 1), ((16, 1, 1), False, 1), ((130, 3, 2), False, 3),
                                               ((23, 4, 3), True, 2)])
def test_single_slab_strict_bit_identical(oracle, hk, scheme, shape, laminar, mrt, idx64):
    w, p, sim = pair(oracle, hk, shape, laminar=laminar, perturb=1e-4, scheme=scheme, idx64=idx64, MRTtype=mrt)
    for step in range(5):                       # both storage phases of the in-place scheme are read back
        w.collision_MRT()
        sim.steps(1)
        assert np.array_equal(sim.download(), w.get_f()), "step %d" % (step + 1)
        w.macrovar()
        for got, k in zip(sim.macrovar(), ("rho", "ux", "uy", "uz")):
            assert np.array_equal(got, w.get(k)), (k, step)
    sim.close()


@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("mrt", [1, 2, 3])
def test_single_slab_fast_arithmetic_within_1e12(oracle, hk, scheme, mrt):
    w, p, sim = pair(oracle, hk, (64, 4, 3), scheme=scheme, strict=False, MRTtype=mrt)
    w.collision_MRT()
    sim.steps(1)
    assert relerr(sim.download(), w.get_f()) < 1e-12          # BASELINE.json: 1 step
    for _ in range(19):
        w.macrovar(); w.collision_MRT()
    sim.steps(19)
    assert relerr(sim.download(), w.get_f()) < 1e-11
    sim.close()


# z-slab runs: the result must not depend on the decomposition, whatever carries the faces
SLABS = [(2, None), (3, None), (3, [2, 2, 3]), (2, [1, 6]), (4, [2, 1, 3, 1])]


@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("nranks,lz", SLABS)
def test_slabs_packed_faces_bit_identical(oracle, hk, scheme, nranks, lz):
    # the NCCL path's kernels: boundary planes (one strided launch) + interior, k_face_pack / k_face_unpack,
    # after an odd in-place step the ghost planes go BACK to the neighbour with the wall-adjacent exclusions
    w, p, sim = pair(oracle, hk, (21, 6, 7), perturb=1e-4, scheme=scheme, nranks=nranks, lz=lz, transport=PACKED)
    assert np.array_equal(sim.download(), w.get_f())
    for step in range(6):
        w.collision_MRT()
        sim.steps(1)
        assert np.array_equal(sim.download(), w.get_f()), "step %d" % (step + 1)
        w.macrovar()
        for got, k in zip(sim.macrovar(), ("rho", "ux", "uy", "uz")):
            assert np.array_equal(got, w.get(k)), (k, step)
    sim.close()


@pytest.mark.parametrize("transport", [FUSED, FUSED_SPLIT, PUT])
@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("nranks,lz", [(2, None), (3, [2, 2, 3]), (3, [3, 2, 2]), (2, [2, 5])])
def test_slabs_peer_memory_halo_bit_identical(oracle, hk, scheme, nranks, lz, transport):
    # halo stored straight into the neighbours' arrays (inside the step kernel, or by the copy-engine transfers) + flag protocol:
    # the harness aborts if a flag or block counter is wrong after a step
    w, p, sim = pair(oracle, hk, (21, 6, 7), perturb=1e-4, scheme=scheme, nranks=nranks, lz=lz, transport=transport)
    for step in range(6):
        w.collision_MRT(); w.macrovar()
        sim.steps(1)
        assert np.array_equal(sim.download(), w.get_f()), "step %d" % (step + 1)
    sim.close()


@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("transport", [PACKED, FUSED, PUT])
def test_slabs_fast_arithmetic_and_64bit_indices(oracle, hk, scheme, transport):
    # production arithmetic is node-local too: FAST on 3 slabs == FAST on one domain, bit for bit; same for the
    # 64-bit-index instantiations of every transport
    shape = (40, 5, 9)
    w, p, one = pair(oracle, hk, shape, perturb=1e-4, scheme=scheme, strict=False)
    _, _, many = pair(oracle, hk, shape, perturb=1e-4, scheme=scheme, strict=False, nranks=3, transport=transport, idx64=True)
    for n in (1, 2, 4):
        one.steps(n); many.steps(n)
        assert np.array_equal(one.download(), many.download()), n
    one.close(); many.close()


@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("nranks", [1, 2])
def test_two_x_blocks_and_pitch_padding(oracle, hk, scheme, nranks):
    # lx = 130: two 128-thread blocks per row, pitch 144 -- 14 padding elements per row that no kernel may touch
    w, p, sim = pair(oracle, hk, (130, 3, 4), perturb=1e-4, scheme=scheme, nranks=nranks)
    for step in range(4):
        w.collision_MRT(); w.macrovar()
        sim.steps(1)
        assert np.array_equal(sim.download(), w.get_f())
    sim.close()


@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("nranks,transport", [(1, PACKED), (3, PACKED), (3, FUSED), (2, FUSED_SPLIT), (3, PUT)])
def test_prerelax_and_external_macro_modes(oracle, hk, scheme, nranks, transport):
    # main.f90:70-90: rhoupdat; collision_MRT with frozen u -- the fused PRERELAX mode of the step kernel; the run-time
    # modes are the GENERIC instantiations, which also exist with the halo stores (pre-relaxation over peer memory)
    nx, ny, nz = 20, 4, 6
    w, p = oracle.make_initial_state(nx, ny, nz, laminar=False, noise=True)
    sim = HostSim(hk, p, nranks=nranks, scheme=scheme, transport=transport)
    sim.upload(w.get_f())
    hk.hs_set_macro(sim.h, *[_d(np.ascontiguousarray(w.get(k))) for k in ("rho", "ux", "uy", "uz")])
    for it in range(5):
        w.rhoupdat()
        w.collision_MRT()
        sim.steps(1, MACRO_PRERELAX)
        assert np.array_equal(sim.download(), w.get_f()), it
    # rhoupdat alone (rho = sum in index order, all nodes)
    w.rhoupdat()
    assert np.array_equal(sim.macrovar(rho_only=True)[0], w.get("rho"))
    # EXTERNAL: conserved moments from arrays the driver set, with a force field
    rng = np.random.default_rng(2)
    shp = (nz, ny, nx)
    F = [1e-5 * rng.normal(size=shp) for _ in range(3)]
    for k, a in zip(("fx", "fy", "fz"), F):
        w.set(k, a)
    hk.hs_set_force_field(sim.h, *[_d(a) for a in F])
    macro = [1e-3 * rng.normal(size=shp)] + [0.02 * rng.normal(size=shp) for _ in range(3)]
    for k, a in zip(("rho", "ux", "uy", "uz"), macro):
        w.set(k, a)
    hk.hs_set_macro(sim.h, *[_d(a) for a in macro])
    for it in range(3):
        w.collision_MRT()
        sim.steps(1, MACRO_EXTERNAL)
        assert np.array_equal(sim.download(), w.get_f()), it
    w.macrovar()
    for got, k in zip(sim.macrovar(), ("rho", "ux", "uy", "uz")):
        assert np.array_equal(got, w.get(k)), k
    w.collision_MRT()
    sim.steps(1, MACRO_MAIN)                    # GENERIC main mode: moments in registers, force from the field
    assert np.array_equal(sim.download(), w.get_f())
    sim.close()


@pytest.mark.parametrize("scheme", [AA, AB])
def test_avedensity_shift_enters_one_collision(oracle, hk, scheme):
    # collision.f90:505-511 shifts the rho ARRAY, not f: the next collision_MRT sees rho - rhomean
    w, p, sim = pair(oracle, hk, (20, 6, 4), perturb=1e-3, scheme=scheme)
    for _ in range(3):
        w.collision_MRT(); w.macrovar()
    sim.steps(3)
    mean, n = w.avedensity()
    hk.hs_set_rho_shift(sim.h, mean)
    w.collision_MRT()
    sim.steps(1)
    assert relerr(sim.download(), w.get_f()) < 1e-15      # rho - mean is formed from a re-summed rho: same bits almost always
    w.macrovar(); w.collision_MRT()
    sim.steps(1)                                           # the shift lived for one collision
    assert relerr(sim.download(), w.get_f()) < 1e-15
    sim.close()


@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("nranks", [1, 2])
def test_solid_nodes_macrovar_and_vortcalc(oracle, hk, scheme, nranks):
    nx, ny, nz = 24, 12, 12
    w, p, sim = pair(oracle, hk, (nx, ny, nz), perturb=1e-4, scheme=scheme, nranks=nranks, ipart=1)
    zz, yy, xx = np.meshgrid(np.arange(nz) + 0.5, np.arange(ny) + 0.5, np.arange(nx) + 0.5, indexing="ij")
    c = (11.3, 1.2, 10.9)                       # straddles the periodic y and z faces
    dy = np.minimum(np.abs(yy - c[1]), ny - np.abs(yy - c[1]))
    dz = np.minimum(np.abs(zz - c[2]), nz - np.abs(zz - c[2]))
    solid = (xx - c[0]) ** 2 + dy ** 2 + dz ** 2 < 3.1 ** 2
    ib = np.where(solid, 1, -1).astype(np.int32)
    isn = np.where(solid, 2, -1).astype(np.int32)
    yp = np.array([[5.0, 5.0, 5.0], list(c)]); wp = np.array([[0, 0, 0], [1e-3, 2e-3, -1e-3]], float)
    om = np.array([[0, 0, 0], [1e-4, -2e-4, 3e-4]], float)
    w.set_solid(ib, isn); w.set_particles(yp, wp, om)
    hk.hs_set_solid(sim.h, _i(ib), _i(isn), 2, _d(yp), _d(wp), _d(om), p.rhopart)
    w.macrovar()
    got = sim.macrovar()
    for g, k in zip(got, ("rho", "ux", "uy", "uz")):
        assert np.array_equal(g, w.get(k)), k
    assert np.all(got[0][solid] == p.rhopart)
    ref = w.vortcalc()
    o = [np.full((nz, ny, nx), np.nan) for _ in range(3)]
    hk.hs_vortcalc(sim.h, *[_d(a) for a in o])
    for a, b in zip(o, ref):
        assert np.array_equal(a, b)
    sim.close()


@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("shape,nranks", [((21, 6, 5), 1), ((21, 6, 5), 2), ((130, 3, 2), 1)])
def test_vortcalc_bit_identical(oracle, hk, scheme, shape, nranks):
    w, p, sim = pair(oracle, hk, shape, perturb=1e-4, scheme=scheme, nranks=nranks)
    w.collision_MRT(); w.macrovar()
    sim.steps(1); sim.macrovar()
    ref = w.vortcalc()
    o = [np.full(shape[::-1], np.nan) for _ in range(3)]
    hk.hs_vortcalc(sim.h, *[_d(a) for a in o])
    for a, b in zip(o, ref):
        assert np.array_equal(a, b)
    sim.close()


@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("nranks", [1, 3])
def test_plane_sums_of_statistc(oracle, hk, scheme, nranks):
    from oracle import oracle as orc
    w, p, sim = pair(oracle, hk, (21, 8, 6), perturb=1e-4, scheme=scheme, nranks=nranks)
    for _ in range(3):
        w.collision_MRT(); w.macrovar()
    sim.steps(3)
    out = np.zeros((12, 21))
    hk.hs_profiles(sim.h, 5, _d(out))
    ref, cnt = orc.plane_sums(w)
    scale = np.abs(ref).max(axis=1, keepdims=True) + 1e-300
    assert np.max(np.abs(out[:11] - ref) / scale) < 1e-13          # order of the (y,z) sum differs
    assert np.array_equal(out[11], cnt)
    sim.close()


@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("nranks", [1, 2])
def test_device_init_matches_initvel_initpop(oracle, hk, scheme, nranks):
    # d3q19_init_channel: log-law + perturbation block (initial.f90:104-144) on the device, un-streamed for AB
    nx, ny, nz = 30, 6, 8
    w, p = oracle.make_initial_state(nx, ny, nz, laminar=False, noise=False, A9=0.3)
    sim = HostSim(hk, p, nranks=nranks, scheme=scheme)
    hk.hs_init_channel(sim.h, p.ustar, p.ystar, 0.3, 0.0, 54321, 1)
    f, ref = sim.download(), w.get_f()
    assert relerr(f, ref) < 1e-14                         # libm sin/cos/exp/log vs the oracle's: same here, not on the device
    # the storage is consistent whatever the scheme: a strict step agrees with the oracle stepping ITS field
    w.set_f(f); w.macrovar(); w.collision_MRT()
    sim.steps(1)
    assert np.array_equal(sim.download(), w.get_f())
    sim.close()


@pytest.mark.parametrize("nranks", [1, 2])
def test_forcingp_against_the_translated_reference(oracle, hk, nranks):
    # FORCINGP (collision.f90:529-602): k_forcingp vs the arrays the translated reference filled at istep = 123
    # (tests/golden/ref_forcingp_*.npz).  Host libm sin/cos here, so the reference's own bits are expected
    import json
    import math
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_forcingp_15x8x8_r2x2_s4.npz"))
    meta = json.loads(str(z["meta"]))
    nx, ny, nz = meta["nx"], meta["ny"], meta["nz"]
    ov = dict(meta["overrides"])
    if "mrttype" in ov:
        ov["MRTtype"] = ov.pop("mrttype")
    if "ustar" in ov:                       # para.f90:64-66: force and y* follow u*
        ov["force_in_y"] = 2.0 * 1.0 * ov["ustar"] * ov["ustar"] / float(nx)
        ov["ystar"] = 0.0036 / ov["ustar"]
    p = oracle.make_para(nx, ny, nz, laminar=meta["laminar"], **ov)
    sim = HostSim(hk, p, nranks=nranks, scheme=AB)
    beta9, gamma9, phase9, ixs0, Tpd = 3.0, 2.0, 0.25, 2, 2000.0                       # :538-545
    pi2 = 2.0 * (4.0 * math.atan(1.0))
    amp0 = 40.00 * beta9 / float(ny) * math.sin(pi2 * float(123) / Tpd)                 # :543
    out = [np.full((nz, ny, nx), np.nan) for _ in range(3)]
    hk.hs_forcingp(sim.h, (nx // 2) // 2, ixs0, p.force_in_y, amp0, beta9, gamma9, phase9, *[_d(a) for a in out])
    scale = float(np.max(np.abs(z["fy"])))
    for a, k in zip(out, ("fx", "fy", "fz")):
        assert np.max(np.abs(a - z[k])) <= 1e-15 * scale, k
    assert np.ptp(out[0]) > 0 and np.ptp(out[2]) > 0
    sim.close()


# ---- kernels that cooperate inside a block (barriers, shared memory, shuffles): one fiber per thread ------------------
@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("nranks", [1, 2])
def test_device_prerelax_loop(oracle, hk, scheme, nranks):
    # d3q19_prerelax: fused rhoupdat + collision, block maximum of |rho - rhop| (shuffles + atomicMax on the bits)
    nx, ny, nz = 32, 8, 8
    tol, maxiter = 2e-5, 12

    def start():
        w, p = oracle.make_initial_state(nx, ny, nz, laminar=False, noise=True)
        w.set_f(w.get_f() + 1e-4 * np.random.default_rng(5).normal(size=(nz, ny, nx, 19)))      # rho no longer relaxed
        return w, p
    w, p = start()
    it_ref = 0
    while True:
        rhop = w.get("rho").copy()
        w.rhoupdat()
        w.collision_MRT()
        err_ref = np.max(np.abs(w.get("rho") - rhop))
        if err_ref <= tol or it_ref > maxiter:                  # main.f90:85
            break
        it_ref += 1
    w0, _ = start()
    sim = HostSim(hk, p, nranks=nranks, scheme=scheme)
    sim.upload(w0.get_f())
    hk.hs_set_macro(sim.h, *[_d(np.ascontiguousarray(w0.get(k))) for k in ("rho", "ux", "uy", "uz")])
    err = C.c_double(0.0)
    it = hk.hs_prerelax(sim.h, tol, maxiter, C.byref(err))
    assert it == it_ref and err.value == err_ref and it >= 2, (it, it_ref, err.value, err_ref)
    assert np.array_equal(sim.download(), w.get_f())
    sim.close()


@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("nranks", [1, 3])
def test_avedensity_reduction(oracle, hk, scheme, nranks):
    w, p, sim = pair(oracle, hk, (20, 6, 6), perturb=1e-3, scheme=scheme, nranks=nranks)
    for _ in range(2):
        w.collision_MRT(); w.macrovar()
    sim.steps(2)
    sim.macrovar()
    mean_ref, n_ref = w.avedensity()
    n = C.c_longlong(0)
    mean = hk.hs_avedensity(sim.h, C.byref(n))
    assert n.value == n_ref == 20 * 6 * 6
    assert abs(mean - mean_ref) <= 1e-13 * np.max(np.abs(w.get_f()))          # the order of the sum differs
    w.collision_MRT()
    sim.steps(1)
    assert relerr(sim.download(), w.get_f()) < 1e-13
    sim.close()


def _diag_from_partials(out, p, nranks, lzs, ustar):
    """host side of d3q19_diag: merge of the per-slab results"""
    nx, ny, nz = p.nx, p.ny, p.nz
    sums = out[:, :7].sum(axis=0)
    vmax, loc = 0.0, (0, 0, 0)
    gz = 0
    for r in range(nranks):
        if out[r, 7] > vmax and out[r, 8] >= 0:
            li = int(out[r, 8])
            vmax, loc = out[r, 7], (li % nx + 1, (li // nx) % ny + 1, li // (nx * ny) + 1 + gz)
        gz += lzs[r]
    nf = sums[0]
    um, vm, wm = sums[1] / nf, sums[2] / nf, sums[3] / nf
    return dict(vmax=vmax, imout=loc[0], jmout=loc[1], kmout=loc[2], umean=um / ustar, vmean=vm / ustar, wmean=wm / ustar,
                urms=np.sqrt(sums[4] / nf - um * um) / ustar, vrms=np.sqrt(sums[5] / nf - vm * vm) / ustar,
                wrms=np.sqrt(sums[6] / nf - wm * wm) / ustar, volf=1.0 - nf / (nx * ny * nz),
                rhomax=out[:, 9].max(), rhomin=out[:, 10].min(), nfluid=int(nf))


@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("nranks", [1, 2])
def test_diag_reduction(oracle, hk, scheme, nranks):
    from oracle import oracle as orc
    nx, ny, nz = 140, 6, 4                     # two x-blocks: the block merge sees two columns sets
    w, p, sim = pair(oracle, hk, (nx, ny, nz), perturb=1e-4, scheme=scheme, nranks=nranks)
    for _ in range(3):
        w.collision_MRT(); w.macrovar()
    sim.steps(3)
    out = np.zeros((nranks, 12))
    hk.hs_diag(sim.h, 5, _d(out))
    lzs = [pkg.slab(nz, nranks, r)[0] for r in range(nranks)]
    got = _diag_from_partials(out, p, nranks, lzs, p.ustar)
    ref = orc.diag_line(w, p.ustar)
    for k in ("imout", "jmout", "kmout", "nfluid"):
        assert got[k] == ref[k], k
    assert got["vmax"] == ref["vmax"] and got["rhomax"] == ref["rhomax"] and got["rhomin"] == ref["rhomin"]
    for k in ("umean", "vmean", "wmean", "urms", "vrms", "wrms"):
        assert abs(got[k] - ref[k]) <= 1e-11 * max(abs(ref[k]), 1e-3), k
    sim.close()


# ---- the particle path against oracle/particles_oracle.c (parity unpinned against the reference: partlib.f90 is absent) ---
PNX, PNY, PNZ, PRAD = 24, 20, 22, 4.3
PPOS = [[11.7, 1.2, 20.9], [5.1, 12.0, 9.0], [15.5, 8.4, 13.2]]      # across the periodic y/z faces, near a wall, in the bulk
PVEL = [[0.010, 0.020, -0.010], [0.0, 0.015, 0.0], [-0.005, 0.0, 0.012]]
POMG = [[1e-3, 0.0, 2e-3], [0.0, -1e-3, 0.0], [5e-4, 5e-4, 0.0]]
PU = dict(ustar=0.0025, ystar=0.0036 / 0.0025, force_in_y=2.0 * 0.0025 * 0.0025 / PNX)


def _particle_pair(oracle, hk, scheme, strict, nranks):
    from oracle import particles as P
    w, p = oracle.make_initial_state(PNX, PNY, PNZ, laminar=False, noise=True, ipart=1, **PU)
    sim = HostSim(hk, p, nranks=nranks, scheme=scheme, strict=strict)
    sim.upload(w.get_f())
    pt = P.Particles(PNX, PNY, PNZ, PRAD, PPOS, PVEL, POMG)
    lub = np.array(pt.lub, dtype=np.float64)
    gf = np.zeros(3)
    hk.hs_particles_init(sim.h, 3, PRAD, 1.0, 1.0, _d(lub), _d(gf), _d(pt.ypglb), _d(pt.wp), _d(pt.omgp))
    return w, p, sim, pt


def _host_links(hk, sim, nranks, cap):
    keys = ("x", "y", "z", "ip", "part")
    parts = []
    for k in range(nranks):
        a = [np.zeros(cap, dtype=np.int32) for _ in range(5)]
        q = np.zeros(cap)
        n = hk.hs_get_links(sim.h, k, *[_i(t) for t in a], _d(q))
        parts.append(dict(zip(keys, [t[:n] for t in a]), q=q[:n]))
    return parts


def _host_particles(hk, sim):
    o = [np.zeros((3, 3)) for _ in range(5)]
    hk.hs_get_particles(sim.h, *[_d(a) for a in o])
    return dict(zip(("ypglb", "wp", "omgp", "fHIp", "torqp"), o))


def _set_oracle_mask(w, pt):
    w.set_solid(np.where(pt.own > 0, 1, -1).astype(np.int32), pt.own)
    w.set_particles(pt.ypglb, pt.wp, pt.omgp)


@pytest.mark.parametrize("scheme", [AA, AB])
@pytest.mark.parametrize("nranks", [1, 2, 3])
def test_particle_mask_and_links_bit_exact(oracle, hk, scheme, nranks):
    w, p, sim, pt = _particle_pair(oracle, hk, scheme, True, nranks)
    own = pt.build_mask()
    k = pt.build_links()
    n = hk.hs_beads_links(sim.h)
    assert n == len(k["q"]) and n > 1000
    got = np.zeros((PNZ, PNY, PNX), dtype=np.int32)
    hk.hs_get_mask(sim.h, _i(got))
    assert np.array_equal(got, own)
    from oracle import particles as P
    parts = _host_links(hk, sim, nranks, n + 8)
    # every slab lists the links whose fluid node it owns; the list is a set, compared after the canonical sort
    lzs = [pkg.slab(PNZ, nranks, r)[0] for r in range(nranks)]
    gz = 0
    for r in range(nranks):
        want, got = P.canon(k, (k["z"] > gz) & (k["z"] <= gz + lzs[r])), P.canon(parts[r])
        for key in ("x", "y", "z", "ip", "part", "q"):
            assert np.array_equal(got[key], want[key]), (r, key)          # same links, same q bits
        gz += lzs[r]
    sim.close()


@pytest.mark.parametrize("scheme,strict,nranks", [(AA, True, 1), (AB, False, 1), (AA, False, 2), (AB, True, 2)])
def test_particle_ibb_and_force(oracle, hk, scheme, strict, nranks):
    w, p, sim, pt = _particle_pair(oracle, hk, scheme, strict, nranks)
    pt.build_mask(); pt.build_links()
    _set_oracle_mask(w, pt)
    w.macrovar()
    fluid = pt.own < 0
    for step in range(4):                                   # both storage phases of AA, fixed particles
        w.collision_MRT()
        f = w.get_f(); pt.ibb(f); w.set_f(f); w.macrovar()
        hk.hs_particle_step(sim.h, 0)
        out = sim.download()
        scale = np.max(np.abs(f[fluid]))
        assert np.max(np.abs(out[fluid] - f[fluid])) < 1e-12 * scale, step
        g = _host_particles(hk, sim)
        assert np.max(np.abs(g["fHIp"] - pt.fHIp)) < 1e-10 * np.max(np.abs(pt.fHIp)), step
        assert np.max(np.abs(g["torqp"] - pt.torqp)) < 1e-10 * np.max(np.abs(pt.torqp)), step
    assert np.dot(pt.fHIp[2], PVEL[2]) < 0                  # drag opposes the motion of the bulk particle
    sim.close()


@pytest.mark.parametrize("scheme,nranks", [(AA, 1), (AB, 3)])
def test_moving_particles_with_refill(oracle, hk, scheme, nranks):
    # z-slabs: a refill next to a face takes its source nodes from the neighbour slab (k_plane_gather + exchange),
    # so the result is that of the single domain
    w, p, sim, pt = _particle_pair(oracle, hk, scheme, False, nranks)
    pt.build_mask(); pt.build_links()
    _set_oracle_mask(w, pt)
    w.macrovar()
    nfill_total = 0
    for step in range(12):
        w.collision_MRT()
        f = w.get_f(); pt.ibb(f)
        pt.lubforce(); pt.move()
        pt.build_mask(); pt.build_links()
        nfill_total += pt.refill(f)
        w.set_f(f); _set_oracle_mask(w, pt); w.macrovar()
        hk.hs_particle_step(sim.h, 1)
        g = _host_particles(hk, sim)
        assert np.max(np.abs(g["ypglb"] - pt.ypglb)) < 1e-11, step
        got = np.zeros((PNZ, PNY, PNX), dtype=np.int32)
        hk.hs_get_mask(sim.h, _i(got))
        assert np.array_equal(got, pt.own), step
        fluid = pt.own < 0
        out = sim.download()
        scale = np.max(np.abs(f[fluid]))
        assert np.max(np.abs(out[fluid] - f[fluid])) < 1e-9 * scale, step
        assert np.max(np.abs(g["fHIp"] - pt.fHIp)) < 1e-9 * np.max(np.abs(pt.fHIp)), step
    assert nfill_total > 0
    if nranks == 1:
        from oracle import particles as P
        k, gl = P.canon(pt.links), P.canon(_host_links(hk, sim, 1, len(pt.links["q"]) + 8)[0])
        for key in ("x", "y", "z", "ip", "part"):             # (q follows the positions, which agree to rounding)
            assert np.array_equal(gl[key], k[key]), key
    sim.close()
