"""Golden vectors produced by RUNNING THE REFERENCE (tests/golden/make_golden.py: the Fortran
hot path machine-translated to C, one thread per MPI rank).  The oracle must reproduce them
bit for bit on CPU; the CUDA path (through the C-ABI) must reproduce them bit for bit in STRICT
arithmetic and within 1e-12 (1 step) / 1e-9 (1000 steps) -- here < 1e-12 over <= 60 steps -- in
FAST arithmetic, on the GPU box where /root/reference does not exist."""
import glob
import json
import os

import numpy as np
import pytest

import __graft_entry__ as entry
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
FIELDS = ("rho", "ux", "uy", "uz")


def load(path):
    z = np.load(path)
    return json.loads(str(z["meta"])), z


def test_fixtures_exist():
    assert len(CASES) >= 11


def _check_stats(z, meta, rows1, rows2, diag, tol=1e-11):
    """profiles.dat / profiles2.dat rows and the diag.dat line against what the reference wrote"""
    for got, want, what in ((rows1, z["profiles"], "profiles"), (rows2, z["profiles2"], "profiles2")):
        if got is None:
            continue
        scale = np.maximum(np.max(np.abs(want), axis=0), 1e-30)
        # the second moments are differences <ab> - <a><b> (saveload.f90:1291-1296): their rounding error is
        # relative to the product of the means, not to the (much smaller) difference
        m = np.max(np.abs(want), axis=0)
        for col, (a, b) in {5: (2, 2), 6: (3, 3), 7: (4, 4), 8: (2, 4), 9: (2, 3), 10: (3, 4), 12: (11, 11)}.items():
            scale[col] = max(scale[col], m[a] * m[b])
        assert np.all(np.abs(got - want) <= tol * scale), (what, float(np.max(np.abs(got - want) / scale)))
    d = z["diag"]
    assert (int(d[2]), int(d[3]), int(d[4])) == (diag["imout"], diag["jmout"], diag["kmout"])
    assert abs(d[1] - diag["vmax"]) <= 1e-14 * d[1]
    assert d[12] == diag["rhomax"] and d[13] == diag["rhomin"]
    for k, v in zip(("umean", "vmean", "wmean", "urms", "vrms", "wrms", "volf"), d[5:12]):
        assert abs(v - diag[k]) <= tol * max(abs(v), 1e-3), k


def oracle_overrides(meta):
    ov = dict(meta["overrides"])
    if "mrttype" in ov:
        ov["MRTtype"] = ov.pop("mrttype")
    if "ustar" in ov:                       # para.f90:64-66: force and y* follow u*
        nx = meta["nx"]
        ov["force_in_y"] = 2.0 * 1.0 * ov["ustar"] * ov["ustar"] / float(nx)
        ov["ystar"] = 0.0036 / ov["ustar"]
    return ov


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
@pytest.mark.parametrize("ranks", ["as_recorded", "single"])
def test_oracle_reproduces_reference_output(path, ranks):
    meta, z = load(path)
    nx, ny, nz = meta["nx"], meta["ny"], meta["nz"]
    npy, npz = meta["ranks"] if ranks == "as_recorded" else (1, 1)
    para = orc.make_para(nx, ny, nz, laminar=meta["laminar"], nprocY=npy, nprocZ=npz, **oracle_overrides(meta))
    for k, v in meta["scalars"].items():
        if k != "mrttype":
            assert getattr(para, k) == v, k
    w = orc.World(para)
    w.FORCING()
    if meta["kind"] == "stats":
        # statistc / statistc2 / diag (saveload.f90:1202-1676) on the state the reference had
        solid = None
        w.set_f(z["f"])
        if meta["solid"]:
            w.close()
            para = orc.make_para(nx, ny, nz, laminar=False, nprocY=npy, nprocZ=npz, ipart=1, **oracle_overrides(meta))
            w = orc.World(para)
            w.FORCING(); w.set_f(z["f"])
            w.set_solid(z["ib"], z["isn"]); w.set_particles(z["ypglb"], z["wp"], z["omgp"])
            solid = z["ib"] > 0
        w.macrovar()
        for k in FIELDS:
            assert np.array_equal(w.get(k), z[k]), k
        rows = entry.load_package().ChannelFlow.statistc_rows
        ustar, ystar = meta["scalars"]["ustar"], meta["scalars"]["ystar"]
        sums_all, _ = orc.plane_sums(w)                     # statistc sums every node, solid ones included (:1241-1266)
        sums_fl, cnt = orc.plane_sums(w, solid)
        _check_stats(z, meta, rows(sums_all, ny * nz, ustar, ystar),
                     rows(sums_fl, cnt, ustar, ystar, with_volf=True, nynz=ny * nz), orc.diag_line(w, ustar, solid))
        w.close()
        return
    w.set_f(z["f0"])
    if meta["kind"] == "prerelax":
        for k in FIELDS:
            w.set(k, z[k + "0"])
        for it in range(meta["steps"]):
            rhop = w.get("rho").copy()
            w.rhoupdat(); w.collision_MRT()
            assert float(np.max(np.abs(w.get("rho") - rhop))) == z["rhoerr"][it]
        assert np.array_equal(w.get("rho"), z["rho"])
    else:
        if meta["kind"] == "force_field":
            for k in ("fx", "fy", "fz"):
                w.set(k, z[k])
        w.macrovar()
        for _ in range(meta["steps"]):
            w.collision_MRT(); w.macrovar()
        for k in FIELDS:
            assert np.array_equal(w.get(k), z[k]), k
        if "ox" in z.files:                              # vortcalc, saveload.f90:3929-4054
            for k, a in zip(("ox", "oy", "oz"), w.vortcalc()):
                assert np.array_equal(a, z[k]), k
        if "sij2" in z.files:                            # sijstat00's strain rate, saveload.f90:2031-2091
            assert np.array_equal(w.sijstat(), z["sij2"])
    assert np.array_equal(w.get_f(), z["f"])
    w.close()


# ---- the CUDA path against the same vectors -------------------------------------------------------
def _sim(meta, scheme, math_mode):
    pkg = entry.load_package()
    ov = dict(meta["overrides"])
    kw = {}
    if "mrttype" in ov:
        kw["MRTtype"] = ov["mrttype"]
    if "ustar" in ov:
        kw.update(ustar=ov["ustar"], force_in_y=2.0 * ov["ustar"] * ov["ustar"] / float(meta["nx"]),
                  ystar=0.0036 / ov["ustar"])
    sim = pkg.ChannelFlow(meta["nx"], meta["ny"], meta["nz"], laminar=meta["laminar"], scheme=scheme,
                          math_mode=math_mode, **kw)
    for k, v in meta["scalars"].items():
        if k != "mrttype":
            assert getattr(sim.v, k) == v, k
    return pkg, sim


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
@pytest.mark.parametrize("scheme", ["aa", "ab"])
@pytest.mark.parametrize("math_mode", ["strict", "fast"])
def test_cuda_path_reproduces_reference_output(path, scheme, math_mode):
    meta, z = load(path)
    pkg = entry.load_package()
    capi = pkg.capi
    pkg, sim = _sim(meta, capi.SCHEME_AA if scheme == "aa" else capi.SCHEME_AB,
                    capi.MATH_STRICT if math_mode == "strict" else capi.MATH_FAST)
    strict = math_mode == "strict"
    if meta["kind"] == "stats":
        # device-side statistc / statistc2 / diag against the numbers the reference wrote to its files
        if meta["solid"]:
            sim.close()
            ov = meta["overrides"]
            sim = pkg.ChannelFlow(meta["nx"], meta["ny"], meta["nz"], laminar=False, ipart=True, math_mode=capi.MATH_STRICT,
                                  scheme=capi.SCHEME_AA if scheme == "aa" else capi.SCHEME_AB, ustar=ov["ustar"],
                                  force_in_y=2.0 * ov["ustar"] * ov["ustar"] / float(meta["nx"]), ystar=0.0036 / ov["ustar"])
        sim.FORCING()
        sim.upload_f(np.ascontiguousarray(z["f"]))
        rows1 = sim.statistc()
        if meta["solid"]:
            sim.set_solid_mask(z["ib"], z["isn"]); sim.set_particles(z["ypglb"], z["wp"], z["omgp"])
            rows1 = None          # statistc sums the rigid-body velocity inside particles too: host arrays, not the device path
        _check_stats(z, meta, rows1, sim.statistc2(), sim.diag())
        sim.close()
        return
    sim.f[...] = z["f0"]
    sim.host_f_changed()
    sim.FORCING()
    scale = max(float(np.max(np.abs(z["f"]))), 1e-300)

    def check(a, b, what):
        if strict:
            assert np.array_equal(a, b), what
        else:
            assert np.max(np.abs(a - b)) < 1e-12 * scale, what

    if meta["kind"] == "prerelax":
        for k in FIELDS:
            getattr(sim, k)[...] = z[k + "0"]
        for it in range(meta["steps"]):                  # the driver's loop, main.f90:70-90
            rhop = sim.rho.copy()
            sim.rhoupdat(); sim.collision_MRT()
            err = float(np.max(np.abs(sim.rho - rhop)))
            if strict:
                assert err == z["rhoerr"][it]
        check(sim.rho, z["rho"], "rho")
    else:
        if meta["kind"] == "force_field":
            sim.set_force_field(np.ascontiguousarray(z["fx"]), np.ascontiguousarray(z["fy"]), np.ascontiguousarray(z["fz"]))
        sim.v.nflowout = 7                               # exercise the download policy too
        sim.run(meta["steps"])                           # collision_MRT; macrovar per step, main.f90:157-161
        for k in FIELDS:                                 # last-step macrovar was downloaded
            check(getattr(sim, k), z[k], k)
        if "ox" in z.files:                              # device vortcalc on the device velocity field
            vscale = max(float(np.max(np.abs(z["oz"]))), 1e-300)
            for k, a in zip(("ox", "oy", "oz"), sim.vortcalc()):
                if strict:
                    assert np.array_equal(a, z[k]), k
                else:
                    assert np.max(np.abs(a - z[k])) < 1e-11 * vscale, k
        if "sij2" in z.files:                            # strain rate from the moments of the device populations
            got, sscale = sim.sijstat(), float(np.max(z["sij2"]))
            if strict:
                assert np.array_equal(got, z["sij2"])
            else:                                        # a difference of nearly equal moments: relative to the largest value
                assert np.max(np.abs(got - z["sij2"])) < 1e-9 * sscale
    check(sim.sync_f_to_host(), z["f"], "f")
    sim.close()


@pytest.mark.gpu
def test_device_forcingp_matches_reference_force_arrays():
    # FORCINGP (collision.f90:529-602) evaluated on the device vs the arrays the translated reference filled
    # at istep = 123 (tests/golden/make_golden.py force_field_case); device sin/cos differ from libm in the last bit
    path = os.path.join(HERE, "golden", "ref_forcingp_15x8x8_r2x2_s4.npz")
    meta, z = load(path)
    pkg = entry.load_package()
    capi = pkg.capi
    pkg, sim = _sim(meta, capi.SCHEME_AB, capi.MATH_FAST)
    sim.FORCINGP(istep=123)
    got = sim.download_force_field()
    scale = float(np.max(np.abs(z["fy"])))
    for a, k in zip(got, ("fx", "fy", "fz")):
        assert np.max(np.abs(a - z[k])) <= 1e-14 * scale, k
    assert np.ptp(got[0]) > 0 and np.ptp(got[2]) > 0
    # and a step with that field agrees with the reference's step to rounding
    sim.f[...] = z["f0"]; sim.host_f_changed()
    sim.run(meta["steps"])
    fscale = float(np.max(np.abs(z["f"])))
    assert np.max(np.abs(sim.sync_f_to_host() - z["f"])) < 1e-12 * fscale
    sim.close()
