"""The compiled-language host above the C-ABI (d3q19-single-phase_b200/host/channel_driver.cpp): main.f90's
call sequence, para, allocarray, initvel and initpop restated in C++, calling the d3q19_shim_* entry points the
Fortran shim binds.

CPU: `--dry-run` (para / initvel / initpop only, no library call) must reproduce the oracle bit for bit.
GPU: the full driver (pre-relaxation loop, time loop with the diag / output cadence, probe) must leave the
populations and macroscopic fields the oracle has after the same sequence -- bit for bit in strict arithmetic,
which also proves that the host arrays were current on every step the driver read them.
"""
import os
import re
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver():
    if os.environ.get("D3Q19_DRIVER"):          # tests/test_hostsim_capi.py: the same driver linked against the host-sim build
        return os.environ["D3Q19_DRIVER"]
    bm = entry._load_build_module()
    return bm.build_driver()


def read_dump(path):
    with open(path, "rb") as fh:
        nx, ny, nz, istep = np.frombuffer(fh.read(16), dtype=np.int32)
        n = int(nx) * int(ny) * int(nz)
        f = np.frombuffer(fh.read(19 * n * 8), dtype=np.float64).reshape(nz, ny, nx, 19)
        fields = [np.frombuffer(fh.read(n * 8), dtype=np.float64).reshape(nz, ny, nx) for _ in range(4)]
    return int(istep), f, dict(zip(("rho", "ux", "uy", "uz"), fields))


def run(driver, *args):
    res = subprocess.run([driver] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    return res.stdout


@pytest.mark.parametrize("shape,laminar,A9", [((64, 32, 32), True, 0.0), ((23, 10, 7), False, 0.3), ((39, 4, 3), False, 0.0),
                                              ((199, 4, 4), False, 0.3)])
def test_dry_run_matches_oracle_para_initvel_initpop(driver, tmp_path, shape, laminar, A9):
    nx, ny, nz = shape
    out = str(tmp_path / "dry.bin")
    args = ["--dry-run", "--nx", nx, "--ny", ny, "--nz", nz, "--A9", A9, "--dump", out] + ([] if laminar else ["--turbulent"])
    text = run(driver, *args)
    w, p = orc.make_initial_state(nx, ny, nz, laminar=laminar, A9=A9, noise=False)
    m = re.search(r"visc (\S+) ustar (\S+) force_in_y (\S+) ystar (\S+) tau (\S+) MRTtype (\d+)", text)
    assert [float(m.group(i)) for i in range(1, 6)] == [p.visc, p.ustar, p.force_in_y, p.ystar, p.tau]
    assert int(m.group(6)) == p.MRTtype
    istep, f, fld = read_dump(out)
    for k in ("ux", "uy", "uz", "rho"):
        assert np.array_equal(fld[k], w.get(k)), k
    assert np.array_equal(f, w.get_f())
    w.close()


@pytest.mark.parametrize("shape,ranks", [((23, 10, 8), 2), ((39, 4, 9), 3), ((16, 3, 7), 3)])
def test_dry_run_ranks_cut_the_channel_like_para(driver, tmp_path, shape, ranks):
    # --ranks N: z-slabs as para.f90:240-262 cuts them (nprocY = 1); the dump is the whole channel again.  The
    # reference's initvel offsets its perturbation by indz*lz (initial.f90:120), exact for even slabs -- the oracle
    # restates that too, so uneven slabs agree with it as well
    nx, ny, nz = shape
    out = str(tmp_path / "dry.bin")
    run(driver, "--dry-run", "--ranks", ranks, "--nx", nx, "--ny", ny, "--nz", nz, "--A9", 0.3, "--turbulent", "--dump", out)
    w, p = orc.make_initial_state(nx, ny, nz, laminar=False, A9=0.3, noise=False, nprocZ=ranks)
    istep, f, fld = read_dump(out)
    for k in ("ux", "uy", "uz", "rho"):
        assert np.array_equal(fld[k], w.get(k)), k
    assert np.array_equal(f, w.get_f())
    w.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ranks", [2, 3])
@pytest.mark.parametrize("scheme", ["aa", "ab"])
def test_driver_ranks_as_threads(driver, tmp_path, scheme, ranks):
    # the N ranks of the reference's job as N threads, one GPU each; pre-relaxation with the all-reduced error, time
    # loop with the diag cadence reduced over ranks, probe of every rank's centre node
    if not os.environ.get("D3Q19_DRIVER") and entry.load_package().capi.device_count() < ranks:
        pytest.skip("needs %d GPUs" % ranks)
    nx, ny, nz, nsteps, itmax = 23, 6, 4 * ranks, 12, 15000
    out = str(tmp_path / "mr.bin")
    U = dict(ustar=0.0025, ystar=0.0036 / 0.0025, force_in_y=2.0 * 0.0025 * 0.0025 / nx)
    text = run(driver, "--ranks", ranks, "--nx", nx, "--ny", ny, "--nz", nz, "--turbulent", "--A9", 0.3, "--ustar", 0.0025,
               "--prerelax", "--nsteps", nsteps, "--ndiag", 5, "--nflowout", 4, "--strict", "--scheme", scheme, "--dump", out)
    w, p = orc.make_initial_state(nx, ny, nz, laminar=False, A9=0.3, noise=False, nprocZ=ranks, **U)
    errs, it = [], 0
    while True:
        rhop = w.get("rho").copy()
        w.rhoupdat(); w.collision_MRT()
        errs.append(float(np.max(np.abs(w.get("rho") - rhop))))
        if errs[-1] <= 1e-5 or it > itmax:
            break
        it += 1
    got = [float(m.group(2)) for m in re.finditer(r"^prerelax (\d+) (\S+)$", text, re.M)]
    assert got == errs
    w.macrovar()
    diag_ref = {}
    for step in range(1, nsteps + 1):
        w.collision_MRT(); w.macrovar()
        if step % 5 == 0:
            diag_ref[step] = orc.diag_line(w, p.ustar)
    istep, f, fld = read_dump(out)
    assert istep == nsteps and np.array_equal(f, w.get_f())
    for k in ("rho", "ux", "uy", "uz"):
        assert np.array_equal(fld[k], w.get(k)), k
    lines = {int(m.group(1)): m.group(2).split() for m in re.finditer(r"^diag (\d+) (.*)$", text, re.M)}
    assert sorted(lines) == [5, 10]
    for step, cols in lines.items():
        assert float(cols[0]) == diag_ref[step]["vmax"]
        assert [int(c) for c in cols[1:4]] == [diag_ref[step][k] for k in ("imout", "jmout", "kmout")]
        assert float(cols[11]) == diag_ref[step]["rhomax"] and float(cols[12]) == diag_ref[step]["rhomin"]
        assert abs(float(cols[5]) - diag_ref[step]["vmean"]) <= 1e-12 * abs(diag_ref[step]["vmean"])
    probes = re.findall(r"^probe(?:_rank \d+)? ", text, re.M)
    assert len(probes) == ranks and len(re.findall(r"^uy_profile ", text, re.M)) == 3
    w.close()


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["aa", "ab"])
def test_driver_main_loop_laminar_config1(driver, tmp_path, scheme):
    # configs[0]: 64x32x32 laminar; nsteps = 60 with diag every 25 and profile output every 20 steps
    nx, ny, nz, nsteps = 64, 32, 32, 60
    out = str(tmp_path / "lam.bin")
    text = run(driver, "--nx", nx, "--ny", ny, "--nz", nz, "--nsteps", nsteps, "--ndiag", 25, "--nflowout", 20, "--strict",
               "--scheme", scheme, "--dump", out)
    w, p = orc.make_initial_state(nx, ny, nz, laminar=True, noise=False)
    w.macrovar()
    diag_ref = {}
    for step in range(1, nsteps + 1):
        w.collision_MRT(); w.macrovar()
        if step % 25 == 0:
            diag_ref[step] = orc.diag_line(w, p.ustar)
    istep, f, fld = read_dump(out)
    assert istep == nsteps
    assert np.array_equal(f, w.get_f())
    for k in ("rho", "ux", "uy", "uz"):
        assert np.array_equal(fld[k], w.get(k)), k
    # the driver's host-side diag read CURRENT host arrays on the diag steps (download policy)
    lines = {int(m.group(1)): m.group(2).split() for m in re.finditer(r"^diag (\d+) (.*)$", text, re.M)}
    assert sorted(lines) == [25, 50]
    for step, cols in lines.items():
        assert float(cols[0]) == diag_ref[step]["vmax"]
        assert [int(c) for c in cols[1:4]] == [diag_ref[step][k] for k in ("imout", "jmout", "kmout")]
        assert float(cols[11]) == diag_ref[step]["rhomax"] and float(cols[12]) == diag_ref[step]["rhomin"]
    pm = re.search(r"^probe (\d+) (\S+) (\S+) (\S+)$", text, re.M)
    c = (nz // 2 - 1, ny // 2 - 1, nx // 2 - 1)
    assert int(pm.group(1)) == nsteps
    assert [float(pm.group(i)) for i in (2, 3, 4)] == [w.get(k)[c] for k in ("ux", "uy", "uz")]
    assert len(re.findall(r"^uy_profile ", text, re.M)) == 3
    w.close()


@pytest.mark.gpu
def test_driver_prerelax_then_main_loop_turbulent(driver, tmp_path):
    # main.f90:70-90 through rhoupdat / collision_MRT with frozen u, then the time loop
    nx, ny, nz, nsteps, itmax = 23, 10, 7, 12, 4
    out = str(tmp_path / "turb.bin")
    text = run(driver, "--nx", nx, "--ny", ny, "--nz", nz, "--turbulent", "--A9", 0.3, "--ustar", 0.0025, "--prerelax",
               "--prerelax-max", itmax, "--nsteps", nsteps, "--ndiag", 5, "--nflowout", 0, "--strict", "--scheme", "aa", "--dump", out)
    U = dict(ustar=0.0025, ystar=0.0036 / 0.0025, force_in_y=2.0 * 0.0025 * 0.0025 / nx)
    w, p = orc.make_initial_state(nx, ny, nz, laminar=False, A9=0.3, noise=False, **U)
    errs, it = [], 0
    while True:
        rhop = w.get("rho").copy()
        w.rhoupdat(); w.collision_MRT()
        errs.append(float(np.max(np.abs(w.get("rho") - rhop))))
        if errs[-1] <= 1e-5 or it > itmax:
            break
        it += 1
    got = [float(m.group(2)) for m in re.finditer(r"^prerelax (\d+) (\S+)$", text, re.M)]
    assert got == errs
    w.macrovar()
    for _ in range(nsteps):
        w.collision_MRT(); w.macrovar()
    istep, f, fld = read_dump(out)
    assert np.array_equal(f, w.get_f())
    for k in ("rho", "ux", "uy", "uz"):
        assert np.array_equal(fld[k], w.get(k)), k
    w.close()
