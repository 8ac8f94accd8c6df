#!/usr/bin/env python
"""Generates tests/golden/*.npz by RUNNING THE REFERENCE: the Fortran hot path of
/root/reference/Channel-Flow machine-translated to C (oracle/f90toc.py -> oracle/_ref/libref.so,
built only in the container where the reference is mounted), one thread per MPI rank.

    python tests/golden/make_golden.py          # needs oracle/_ref/libref.so (make -C oracle ref)

Every fixture stores the inputs (initial populations f0, frozen u or force arrays where the
case needs them) and the reference's outputs after `steps` steps of the driver sequence named
in `kind`, plus the parameters as a JSON string.  Consumers: tests/test_golden.py (the oracle
on CPU; the CUDA path through the C-ABI on the GPU box, where /root/reference does not exist).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS = ("rho", "ux", "uy", "uz")


def noise(shape, ustar, seed):
    rng = np.random.default_rng(seed)
    return [1e-3 * ustar * (2.0 * rng.random(shape) - 1.0) for _ in range(3)]


def start(nx, ny, nz, npy, npz, laminar, seed=None, a9=0.0, **ov):
    w = ref.RefWorld(nx, ny, nz, nprocY=npy, nprocZ=npz, laminar=laminar, a9=a9, **ov)
    w.run("initvel")                                   # main.f90:58
    if seed is not None:
        for k, d in zip(("ux", "uy", "uz"), noise((nz, ny, nx), w.scalar("ustar"), seed)):
            w.set(k, w.get(k) + d)
    w.run("forcing")                                   # main.f90:61
    w.run("initpop")                                   # main.f90:65
    return w


def scalars(w):
    keys = ("visc", "ustar", "ystar", "force_in_y", "force_mag", "s1", "s2", "s4", "s9", "s10", "s13", "s16",
            "omegepsl", "omegepslj", "omegxx")
    d = {k: float(w.scalar(k)) for k in keys}
    d["mrttype"] = int(w.scalar("mrttype"))
    return d


def save(name, meta, **arrays):
    meta = dict(meta)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta), **arrays)
    print("wrote %s.npz: %s" % (name, {k: v.shape for k, v in arrays.items()}))


def main_loop_case(name, nx, ny, nz, npy, npz, laminar, steps, seed, a9=0.0, vort=False, snap=None, sij=False, **ov):
    w = start(nx, ny, nz, npy, npz, laminar, seed, a9, **ov)
    f0 = w.get_f().copy()
    w.run("macrovar")                                  # main.f90:136
    extra = {}
    for it in range(1, steps + 1):
        w.run("collision_mrt")                         # main.f90:157
        w.run("macrovar")                              # main.f90:161
        if snap is not None and it == snap:            # an intermediate state (the other in-place storage phase)
            extra["f_snap"] = w.get_f().copy()
    if vort:
        w.run("vortcalc")                              # saveload.f90:3929 (as called by outputvort1, :1138)
        extra.update({k: w.get(k) for k in ("ox", "oy", "oz")})
    if sij:
        extra["sij2"] = w.sij2()                       # saveload.f90:2031-2091 (first loop nest of sijstat00)
    meta = dict(kind="main_loop", nx=nx, ny=ny, nz=nz, ranks=[npy, npz], laminar=laminar, steps=steps,
                overrides=ov, scalars=scalars(w))
    if snap is not None:
        meta["snap"] = snap
    save(name, meta, f0=f0, f=w.get_f(), **extra, **{k: w.get(k) for k in FIELDS})
    w.close()


def prerelax_case(name, nx, ny, nz, npy, npz, iters, seed, **ov):
    w = start(nx, ny, nz, npy, npz, False, seed, **ov)
    f0 = w.get_f().copy()
    u0 = {k + "0": w.get(k).copy() for k in FIELDS}
    errs = []
    for _ in range(iters):                             # main.f90:70-90
        rhop = w.get("rho").copy()
        w.run("rhoupdat")
        w.run("collision_mrt")
        errs.append(float(np.max(np.abs(w.get("rho") - rhop))))
    meta = dict(kind="prerelax", nx=nx, ny=ny, nz=nz, ranks=[npy, npz], laminar=False, steps=iters, overrides=ov,
                scalars=scalars(w))
    save(name, meta, f0=f0, f=w.get_f(), rho=w.get("rho"), rhoerr=np.array(errs), **u0)
    w.close()


def force_field_case(name, nx, ny, nz, npy, npz, steps, seed, **ov):
    w = start(nx, ny, nz, npy, npz, False, seed, **ov)
    f0 = w.get_f().copy()
    w.set_scalar("istep", 123)
    w.run("forcingp")                                  # collision.f90:529-602
    force = {k: w.get("force_real" + k[-1]).copy() for k in ("fx", "fy", "fz")}
    w.run("macrovar")
    for _ in range(steps):
        w.run("collision_mrt")
        w.run("macrovar")
    meta = dict(kind="force_field", nx=nx, ny=ny, nz=nz, ranks=[npy, npz], laminar=False, steps=steps, overrides=ov,
                scalars=scalars(w))
    save(name, meta, f0=f0, f=w.get_f(), **force, **{k: w.get(k) for k in FIELDS})
    w.close()


def stats_case(name, nx, ny, nz, npy, npz, steps, seed, solid=False, a9=0.3, **ov):
    """profiles.dat / profiles2.dat / diag.dat numbers (statistc, statistc2, diag: saveload.f90:1202-1676) as the
    reference writes them after `steps` steps; with solid=True a sphere is marked solid (ibnodes > 0, owned by
    particle 2) in the INITIAL state and the monitors run right after macrovar -- the reference ships no
    particle code, so a field evolved next to solid nodes is not defined by it."""
    w = start(nx, ny, nz, npy, npz, False, seed, a9, **({"ipart": True} if solid else {}), **ov)
    extra = {}
    if solid:
        zz, yy, xx = np.meshgrid(np.arange(nz) + 0.5, np.arange(ny) + 0.5, np.arange(nx) + 0.5, indexing="ij")
        c = np.array([nx * 0.45, ny * 0.5 + 0.3, nz * 0.5 - 0.2])
        mask = (xx - c[0]) ** 2 + (yy - c[1]) ** 2 + (zz - c[2]) ** 2 < (0.3 * min(nx, ny, nz)) ** 2
        ib = np.where(mask, 1, -1).astype(np.int32)
        isn = np.where(mask, 2, -1).astype(np.int32)
        npart = w.array("ypglb")[0].shape[0]
        yp = np.zeros((npart, 3)); wp = np.zeros((npart, 3)); om = np.zeros((npart, 3))
        yp[1], wp[1], om[1] = c, [0.004, -0.002, 0.001], [1e-3, 2e-3, -1e-3]
        w.set_solid(ib, isn)
        for r in range(w.nproc):
            for nm, a in (("ypglb", yp), ("wp", wp), ("omgp", om)):
                w.array(nm, r)[0][...] = a
        extra = dict(ib=ib, isn=isn, ypglb=yp, wp=wp, omgp=om)
        steps = 0
    w.run("macrovar")
    for _ in range(steps):
        w.run("collision_mrt")
        w.run("macrovar")
    w.set_scalar("istep", steps)
    w.clear_captured()
    w.run("statistc"); w.run("statistc2"); w.run("diag")
    u27 = w.captured(27)
    rows1 = u27[1:1 + 13 * nx].reshape(nx, 13)
    rows2 = u27[2 + 13 * nx:].reshape(nx, 15)
    meta = dict(kind="stats", nx=nx, ny=ny, nz=nz, ranks=[npy, npz], laminar=False, steps=steps, overrides=ov,
                scalars=scalars(w), solid=bool(solid))
    save(name, meta, f=w.get_f(), profiles=rows1, profiles2=rows2, diag=w.captured(26), **extra,
         **{k: w.get(k) for k in FIELDS})
    w.close()


if __name__ == "__main__":
    if not ref.available():
        raise SystemExit("oracle/_ref/libref.so missing: run `make -C oracle ref` where /root/reference is mounted")
    # The turbulent set's u* = 2 Re_tau nu / nx (para.f90:64) is meant for nx ~ 200-500; on these
    # tiny channels it is replaced by configs[1]'s wall units (u* = 0.0025) so the flow is stable.
    U = dict(ustar=0.0025)
    main_loop_case("ref_turb_mrt1_7x8x8_r2x2_s20", 7, 8, 8, 2, 2, False, 20, seed=54321, a9=0.3, **U)
    main_loop_case("ref_lam_lbgk_7x8x8_r1x2_s60", 7, 8, 8, 1, 2, True, 60, seed=None)
    main_loop_case("ref_turb_mrt3_9x10x7_r3x2_s10", 9, 10, 7, 3, 2, False, 10, seed=777, mrttype=3, **U)
    main_loop_case("ref_turb_mrt1_39x4x3_r1x1_s8", 39, 4, 3, 1, 1, False, 8, seed=4242, **U)   # reaches the log-law branch
    main_loop_case("ref_turb_vort_21x6x5_r2x2_s6", 21, 6, 5, 2, 2, False, 6, seed=31337, a9=0.3, vort=True, **U)
    # z-slab parity case of bench.py (world > 1) and tests/test_gpu_multi.py: 16 planes cut into 2 / 4 / 8 (or 3, 5: uneven)
    # slabs, 9 steps (odd: the in-place scheme ends in its swapped phase and has sent its ghost planes back four times),
    # the state after 8 steps stored as well
    main_loop_case("ref_slabs_21x4x16_r1x4_s9", 21, 4, 16, 1, 4, False, 9, seed=8086, a9=0.3, snap=8, **U)
    main_loop_case("ref_turb_sij_13x6x7_r1x2_s5", 13, 6, 7, 1, 2, False, 5, seed=1618, a9=0.3, sij=True, **U)
    stats_case("ref_stats_21x8x6_r2x2_s6", 21, 8, 6, 2, 2, 6, seed=2024, **U)
    stats_case("ref_stats_solid_24x12x12_r1x2", 24, 12, 12, 1, 2, 0, seed=11, solid=True, **U)
    prerelax_case("ref_prerelax_7x8x8_r2x2_i6", 7, 8, 8, 2, 2, 6, seed=99, **U)
    force_field_case("ref_forcingp_15x8x8_r2x2_s4", 15, 8, 8, 2, 2, 4, seed=5, **U)
