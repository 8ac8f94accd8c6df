"""Pins the hand-written oracle (oracle/d3q19_oracle.c) to the reference ITSELF: the
reference's own Fortran hot path machine-translated to C at build time (oracle/f90toc.py ->
oracle/_ref/libref.so, built only where /root/reference is mounted) and run here with one
thread per MPI rank.  Agreement must be bit for bit: both are IEEE evaluations of the same
expression order (-ffp-contract=off).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref.so not built (no /root/reference here)")

FIELDS = ("rho", "ux", "uy", "uz")


def pair(nx, ny, nz, npy, npz, laminar, noise=True, A9=0.0, **ov):
    rw = ref.RefWorld(nx, ny, nz, nprocY=npy, nprocZ=npz, laminar=laminar, a9=A9, **ov)
    oov = {{"mrttype": "MRTtype"}.get(k, k): v for k, v in ov.items()}
    para = orc.make_para(nx, ny, nz, laminar=laminar, nprocY=npy, nprocZ=npz, **oov)
    ow = orc.World(para)
    rw.run("initvel")
    ow.initvel(A9)
    for k in ("ux", "uy", "uz"):
        assert np.array_equal(rw.get(k), ow.get(k)), "initvel " + k
    if noise:
        for k, d in zip(("ux", "uy", "uz"), orc.synthetic_velocity(nx, ny, nz, para.ustar)):
            rw.set(k, rw.get(k) + d)
            ow.set(k, ow.get(k) + d)
    rw.run("forcing"); ow.FORCING()
    rw.run("initpop"); ow.initpop()
    assert np.array_equal(rw.get_f(), ow.get_f()), "initpop"
    return rw, ow, para


def same(rw, ow, what=""):
    assert np.array_equal(rw.get_f(), ow.get_f()), "f " + what
    for k in FIELDS:
        assert np.array_equal(rw.get(k), ow.get(k)), k + " " + what


def test_para_constants_match():
    for laminar in (True, False):
        rw = ref.RefWorld(199, 200, 200, laminar=laminar)     # the shipped size, var_inc.f90:51
        p = orc.make_para(199, 200, 200, laminar=laminar)
        for k in ("visc", "ustar", "ystar", "force_in_y", "tau", "s1", "s2", "s4", "s9", "s10", "s13", "s16",
                  "omegepsl", "omegepslj", "omegxx", "coef1", "coef3i", "coef4i", "val2i", "val3i", "val9i",
                  "ww0", "ww1", "ww2", "pi", "pi2"):
            assert rw.scalar(k) == getattr(p, k), k
        assert rw.scalar("mrttype") == p.MRTtype
        for k in ("cix", "ciy", "ciz", "ipopp", "ipswap", "ipstay"):
            assert list(rw.array(k)[0]) == list(getattr(p, k)), k
        rw.close()


@pytest.mark.parametrize("npy,npz", [(1, 1), (2, 2), (1, 2), (2, 1), (1, 4), (3, 2)])
@pytest.mark.parametrize("laminar", [False, True])
def test_main_loop_bit_exact(npy, npz, laminar):
    # the reference's own size tie nx = nx7-1, ny = nz = nx7 (var_inc.f90:51) at nx7 = 12
    rw, ow, p = pair(11, 12, 12, npy, npz, laminar)
    rw.run("macrovar"); ow.macrovar()
    same(rw, ow, "after first macrovar")
    for step in range(12):
        rw.run("collision_mrt"); rw.run("macrovar")      # main.f90:157-161
        ow.collision_MRT(); ow.macrovar()
        same(rw, ow, "step %d" % step)
    rw.close(); ow.close()


def test_uneven_split_and_free_sizes():
    rw, ow, p = pair(9, 10, 7, 3, 2, False)               # ly = 4,3,3  lz = 4,3 (para.f90:233-245)
    assert [rw.scalar("ly", r) for r in range(6)] == [4, 3, 3, 4, 3, 3]
    assert [rw.scalar("lz", r) for r in range(6)] == [4, 4, 4, 3, 3, 3]
    rw.run("macrovar"); ow.macrovar()
    for step in range(8):
        rw.run("collision_mrt"); rw.run("macrovar")
        ow.collision_MRT(); ow.macrovar()
    same(rw, ow)
    rw.close(); ow.close()


def test_decomposition_invariance_of_the_reference():
    # SURVEY fact 8, checked on the reference itself: the rank grid does not change a single bit
    res = []
    for npy, npz in [(1, 1), (2, 3)]:
        rw, ow, p = pair(7, 8, 9, npy, npz, False)
        rw.run("macrovar")
        for _ in range(6):
            rw.run("collision_mrt"); rw.run("macrovar")
        res.append(rw.get_f())
        rw.close(); ow.close()
    assert np.array_equal(res[0], res[1])


@pytest.mark.parametrize("mrt", [1, 2, 3])
def test_mrt_types(mrt):
    rw, ow, p = pair(7, 8, 8, 2, 1, False, mrttype=mrt)
    assert rw.scalar("s1") == p.s1 and rw.scalar("omegxx") == p.omegxx
    rw.run("macrovar"); ow.macrovar()
    for _ in range(5):
        rw.run("collision_mrt"); rw.run("macrovar")
        ow.collision_MRT(); ow.macrovar()
    same(rw, ow)
    rw.close(); ow.close()


def test_initvel_perturbation_block():
    # A9 is hard-wired to 0.0 (initial.f90:84); the override exercises the block it disables
    rw, ow, p = pair(15, 16, 16, 2, 2, False, noise=False, A9=0.3)
    assert np.max(np.abs(rw.get("ux"))) > 0
    rw.close(); ow.close()


def test_prerelaxation_sequence():
    # main.f90:70-90: rhop = rho; rhoupdat; collision_MRT with u frozen
    rw, ow, p = pair(11, 12, 12, 2, 2, False)
    for it in range(6):
        rw.run("rhoupdat"); ow.rhoupdat()
        rw.run("collision_mrt"); ow.collision_MRT()
        same(rw, ow, "prerelax %d" % it)
    rw.close(); ow.close()


def test_avedensity():
    rw, ow, p = pair(7, 8, 8, 2, 2, False)
    rw.run("macrovar"); ow.macrovar()
    for _ in range(3):
        rw.run("collision_mrt"); rw.run("macrovar")
        ow.collision_MRT(); ow.macrovar()
    rw.run("avedensity")
    mean, n = ow.avedensity()
    assert n == 7 * 8 * 8
    # the global sum is order dependent (rank partial sums): tolerance, not bits
    assert np.allclose(rw.get("rho"), ow.get("rho"), rtol=0, atol=1e-15 * np.max(np.abs(ow.get_f())))
    rw.close(); ow.close()


def test_force_field_path_with_forcingp():
    # FORCINGP has no live call site (main.f90:133,149-154) but defines the array-force path
    rw, ow, p = pair(15, 16, 16, 2, 2, False)
    rw.set_scalar("istep", 123)
    rw.run("forcingp")
    for k, name in (("fx", "force_realx"), ("fy", "force_realy"), ("fz", "force_realz")):
        ow.set(k, rw.get(name))
    assert np.ptp(rw.get("force_realx")) > 0
    rw.run("macrovar"); ow.macrovar()
    for _ in range(4):
        rw.run("collision_mrt"); rw.run("macrovar")
        ow.collision_MRT(); ow.macrovar()
    same(rw, ow)
    rw.close(); ow.close()


def test_solid_nodes_and_macrovar_solid_branch():
    # ibnodes > 0: collision.f90:54 skips the collision and the stay-loop but still runs the
    # swap-loop with the previous node's f9/Fbar; macrovar's solid branch (collision.f90:420-459)
    nx, ny, nz = 11, 12, 12
    rw = ref.RefWorld(nx, ny, nz, nprocY=1, nprocZ=2, laminar=False, ipart=True)
    para = orc.make_para(nx, ny, nz, laminar=False, nprocY=1, nprocZ=2, ipart=1)
    ow = orc.World(para)
    rw.run("initvel"); ow.initvel(0.0)
    for k, d in zip(("ux", "uy", "uz"), orc.synthetic_velocity(nx, ny, nz, para.ustar)):
        rw.set(k, rw.get(k) + d); ow.set(k, ow.get(k) + d)
    rw.run("forcing"); ow.FORCING()
    rw.run("initpop"); ow.initpop()
    # one sphere of radius 2.6 centred in the channel, owned by particle 3
    zz, yy, xx = np.meshgrid(np.arange(nz) + 0.5, np.arange(ny) + 0.5, np.arange(nx) + 0.5, indexing="ij")
    c = np.array([5.3, 6.1, 6.4])
    solid = (xx - c[0]) ** 2 + (yy - c[1]) ** 2 + (zz - c[2]) ** 2 < 2.6 ** 2
    ib = np.where(solid, 1, -1).astype(np.int32)
    isn = np.where(solid, 3, -1).astype(np.int32)
    npart = rw.array("ypglb")[0].shape[0]
    yp = np.zeros((npart, 3)); wp = np.zeros((npart, 3)); om = np.zeros((npart, 3))
    yp[2], wp[2], om[2] = c, [0.01, -0.02, 0.005], [1e-3, 2e-3, -1e-3]
    rw.set_solid(ib, isn); ow.set_solid(ib, isn)
    for r in range(rw.nproc):
        for name, a in (("ypglb", yp), ("wp", wp), ("omgp", om)):
            rw.array(name, r)[0][...] = a
    ow.set_particles(yp, wp, om)
    rw.run("macrovar"); ow.macrovar()
    same(rw, ow, "solid macrovar")
    assert np.any(rw.get("rho") == 1.0)            # rhopart written on solid nodes
    for step in range(4):
        rw.run("collision_mrt"); rw.run("macrovar")
        ow.collision_MRT(); ow.macrovar()
        same(rw, ow, "solid step %d" % step)
    rw.close(); ow.close()


def _vort_equal(rw, ow, what):
    rw.run("vortcalc")
    got = ow.vortcalc()
    for k, g in zip(("ox", "oy", "oz"), got):
        assert np.array_equal(rw.get(k), g), "%s %s" % (k, what)
    return got


@pytest.mark.parametrize("npy,npz", [(1, 1), (1, 2), (2, 2), (3, 2), (1, 4)])
def test_vortcalc_bit_exact(npy, npz):
    # saveload.f90:3929-4054: vorticity by central differences, one-sided at the walls, neighbour planes
    # through exchng8 -- the translated reference (array-section arguments copied in and out) vs the restatement
    rw, ow, p = pair(9, 12, 8, npy, npz, False, A9=0.3)
    rw.run("macrovar"); ow.macrovar()
    for step in range(3):
        rw.run("collision_mrt"); rw.run("macrovar")
        ow.collision_MRT(); ow.macrovar()
    ox, oy, oz = _vort_equal(rw, ow, "grid %dx%d" % (npy, npz))
    assert np.max(np.abs(oz)) > 0 and np.max(np.abs(ox)) > 0
    # the result does not depend on the decomposition
    p1 = orc.make_para(9, 12, 8, laminar=False)
    o1 = orc.World(p1)
    for k in ("ux", "uy", "uz"):
        o1.set(k, ow.get(k))
    for a, b in zip(o1.vortcalc(), (ox, oy, oz)):
        assert np.array_equal(a, b)
    rw.close(); ow.close(); o1.close()


def test_vortcalc_solid_nodes():
    # inside a particle the vorticity is twice the particle's angular velocity (saveload.f90:4008-4019)
    nx, ny, nz = 11, 12, 12
    rw = ref.RefWorld(nx, ny, nz, nprocY=1, nprocZ=2, laminar=False, ipart=True)
    para = orc.make_para(nx, ny, nz, laminar=False, nprocY=1, nprocZ=2, ipart=1)
    ow = orc.World(para)
    rw.run("initvel"); ow.initvel(0.0)
    for k, d in zip(("ux", "uy", "uz"), orc.synthetic_velocity(nx, ny, nz, para.ustar)):
        rw.set(k, rw.get(k) + d); ow.set(k, ow.get(k) + d)
    rw.run("forcing"); ow.FORCING()
    rw.run("initpop"); ow.initpop()
    zz, yy, xx = np.meshgrid(np.arange(nz) + 0.5, np.arange(ny) + 0.5, np.arange(nx) + 0.5, indexing="ij")
    c = np.array([5.3, 6.1, 6.4])
    solid = (xx - c[0]) ** 2 + (yy - c[1]) ** 2 + (zz - c[2]) ** 2 < 2.6 ** 2
    ib = np.where(solid, 1, -1).astype(np.int32)
    isn = np.where(solid, 3, -1).astype(np.int32)
    npart = rw.array("ypglb")[0].shape[0]
    yp = np.zeros((npart, 3)); wp = np.zeros((npart, 3)); om = np.zeros((npart, 3))
    yp[2], wp[2], om[2] = c, [0.01, -0.02, 0.005], [1e-3, 2e-3, -1e-3]
    rw.set_solid(ib, isn); ow.set_solid(ib, isn)
    for r in range(rw.nproc):
        for name, a in (("ypglb", yp), ("wp", wp), ("omgp", om)):
            rw.array(name, r)[0][...] = a
    ow.set_particles(yp, wp, om)
    rw.run("macrovar"); ow.macrovar()
    ox, oy, oz = _vort_equal(rw, ow, "solid")
    assert np.all(ox[solid] == 2e-3) and np.all(oy[solid] == 4e-3) and np.all(oz[solid] == -2e-3)
    rw.close(); ow.close()


# ---- saveload.f90 monitors: the numbers the translated reference writes to its files ----------------------
def _stat_rows():
    import __graft_entry__ as entry
    return entry.load_package().ChannelFlow.statistc_rows


@pytest.mark.parametrize("npy,npz", [(1, 1), (2, 2), (1, 4)])
def test_statistc_and_diag_file_output(npy, npz):
    # profiles.dat (statistc, saveload.f90:1202-1342), profiles2.dat (statistc2, :1348-1502) and diag.dat
    # (diag, :1507-1676) as the reference writes them, vs the numpy restatements the GPU tests check the
    # device reductions against, and the host-side normalisation of the plane sums (channel.py statistc_rows)
    nx, ny, nz = 11, 12, 12
    rw, ow, p = pair(nx, ny, nz, npy, npz, False, A9=0.3)
    rw.run("macrovar"); ow.macrovar()
    for step in range(5):
        rw.run("collision_mrt"); rw.run("macrovar")
        ow.collision_MRT(); ow.macrovar()
    rw.set_scalar("istep", 5)
    rw.clear_captured()
    rw.run("statistc"); rw.run("statistc2"); rw.run("diag")
    ustar, ystar = rw.scalar("ustar"), rw.scalar("ystar")
    out27 = rw.captured(27)
    # unit 27 holds: istep, lx rows of 13 (statistc), istep, lx rows of 15 (statistc2)
    assert out27.size == 1 + 13 * nx + 1 + 15 * nx and out27[0] == 5 and out27[1 + 13 * nx] == 5
    rows1 = out27[1:1 + 13 * nx].reshape(nx, 13)
    rows2 = out27[2 + 13 * nx:].reshape(nx, 15)
    sums, cnt = orc.plane_sums(ow)
    mine1 = _stat_rows()(sums, ny * nz, ustar, ystar)
    mine2 = _stat_rows()(sums, cnt, ustar, ystar, with_volf=True, nynz=ny * nz)
    for got, want in ((mine1, rows1), (mine2, rows2)):
        scale = np.maximum(np.max(np.abs(want), axis=0), 1e-30)
        m = np.max(np.abs(want), axis=0)       # second moments are differences <ab> - <a><b>: error relative to <a><b>
        for col, (a, b) in {5: (2, 2), 6: (3, 3), 7: (4, 4), 8: (2, 4), 9: (2, 3), 10: (3, 4), 12: (11, 11)}.items():
            scale[col] = max(scale[col], m[a] * m[b])
        assert np.all(np.abs(got - want) <= 1e-11 * scale), np.max(np.abs(got - want) / scale)
    assert np.array_equal(mine1[:, :2], rows1[:, :2])                   # x and y+ columns are exact
    assert np.all(rows2[:, 13] == 1.0) and np.all(rows2[:, 14] == 0.0)  # no solids: volume fractions
    d = rw.captured(26)
    assert d.size == 14
    mine = orc.diag_line(ow, ustar)
    assert d[0] == 5.0
    assert (int(d[2]), int(d[3]), int(d[4])) == (mine["imout"], mine["jmout"], mine["kmout"])
    assert d[1] == mine["vmax"] and d[12] == mine["rhomax"] and d[13] == mine["rhomin"]
    for k, v in zip(("umean", "vmean", "wmean", "urms", "vrms", "wrms", "volf"), d[5:12]):
        assert abs(v - mine[k]) <= 1e-11 * max(abs(v), 1e-3), k
    rw.close(); ow.close()


@pytest.mark.parametrize("npy,npz,mrt", [(1, 1, 1), (2, 2, 1), (1, 3, 3), (3, 2, 2)])
def test_sijstat_strain_rate_bit_exact(npy, npz, mrt):
    # saveload.f90:2031-2091 (first loop nest of sijstat00): Sij*Sij from the non-equilibrium moments.  The translated
    # routine keeps it in an automatic array whose elements are captured; the restatement must give the same bits.
    rw, ow, p = pair(9, 12, 9, npy, npz, False, A9=0.3, mrttype=mrt)
    rw.run("macrovar"); ow.macrovar()
    for step in range(3):
        rw.run("collision_mrt"); rw.run("macrovar")
        ow.collision_MRT(); ow.macrovar()
    a, b = rw.sij2(), ow.sijstat()
    assert np.array_equal(a, b)
    assert np.min(a) >= 0.0 and np.max(a) > 0.0 and np.isfinite(a).all()
    rw.close(); ow.close()


def test_sijstat_skips_solid_nodes():
    nx, ny, nz = 11, 12, 12
    rw = ref.RefWorld(nx, ny, nz, nprocY=1, nprocZ=2, laminar=False, ipart=True)
    para = orc.make_para(nx, ny, nz, laminar=False, nprocY=1, nprocZ=2, ipart=1)
    ow = orc.World(para)
    rw.run("initvel"); ow.initvel(0.0)
    for k, d in zip(("ux", "uy", "uz"), orc.synthetic_velocity(nx, ny, nz, para.ustar)):
        rw.set(k, rw.get(k) + d); ow.set(k, ow.get(k) + d)
    rw.run("forcing"); ow.FORCING()
    rw.run("initpop"); ow.initpop()
    zz, yy, xx = np.meshgrid(np.arange(nz) + 0.5, np.arange(ny) + 0.5, np.arange(nx) + 0.5, indexing="ij")
    mask = (xx - 5.2) ** 2 + (yy - 6.1) ** 2 + (zz - 6.3) ** 2 < 9.0
    ib, isn = np.where(mask, 1, -1).astype(np.int32), np.where(mask, 1, -1).astype(np.int32)
    rw.set_solid(ib, isn); ow.set_solid(ib, isn)
    a, b = rw.sij2(), ow.sijstat()
    assert np.array_equal(a[~mask], b[~mask]) and np.all(b[mask] == 0.0)
    rw.close(); ow.close()
