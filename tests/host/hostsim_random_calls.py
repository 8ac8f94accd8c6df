"""TEST INFRASTRUCTURE ONLY.  Random sequences of C-ABI calls on the host-sim build of the library, mirrored on the oracle
and compared after every call that reads state: steps in bursts (so both in-place storage phases are met by every
reader), downloads, macrovar, probe, plane sums, re-uploads in either storage phase, uniform force / force field
switches, frozen-moment (EXTERNAL) and PRERELAX steps, avedensity with the shifted collision that follows.  Strict
arithmetic: everything but the order-dependent sums must agree bit for bit.

    D3Q19_LIB=tests/host/_gen/libd3q19b200_hostsim.so python tests/host/hostsim_random_calls.py [nseeds]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
from oracle import oracle as orc  # noqa: E402

pkg = entry.load_package()
capi = pkg.capi


def run_sequence(seed, scheme, log, rank=0, world=1, comm=None, transport="nccl"):
    """raw C-ABI calls in random order; with world > 1 every rank (a thread, tests/host/hostsim_mrank_worker.py) follows
    the same seeded sequence on its z-slab, over the given halo transport"""
    rng = np.random.default_rng(seed)
    nx, ny = int(rng.integers(5, 40)), int(rng.integers(1, 6))
    nz = int(rng.integers(1, 6)) if world == 1 else int(rng.integers(2 * world, 3 * world + 3))
    U = dict(ustar=0.0025, ystar=0.0036 / 0.0025, force_in_y=2.0 * 0.0025 * 0.0025 / nx)
    w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True, **U)
    w.set_f(w.get_f() + 1e-4 * rng.normal(size=(nz, ny, nx, 19)))
    w.macrovar()
    # HOSTSIM_FAST=1: production arithmetic; agreement to rounding instead of bit for bit
    fast = bool(os.environ.get("HOSTSIM_FAST"))
    sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, scheme=scheme, math_mode=capi.MATH_FAST if fast else capi.MATH_STRICT,
                          rank=rank, nranks=world, device=0, nccl_id=comm.new_id(rank) if world > 1 else None,
                          halo_split_min=3 if transport == "peer-split" else 0, **U)
    sim.FORCING()
    if transport in ("peer", "peer-split", "put"):
        assert sim.connect_halo(lambda b: comm.allgather(rank, bytes(b)), mode="put" if transport.startswith("put") else "fused")
    sl = slice(sim.globalz, sim.globalz + sim.lz)
    sim.upload_f(np.ascontiguousarray(w.get_f()[sl]))
    shp, lshp = (nz, ny, nx), (sim.lz, ny, nx)
    out = np.empty(lshp + (19,))
    exact = not fast                  # False between an avedensity and the next re-upload (order-dependent mean)
    rtol = 1e-11 if fast else 1e-13
    ops = ["step", "step", "step", "download", "macro", "probe", "sums", "reupload", "force", "field", "external", "prerelax",
           "avedensity"]
    for k in range(40 if world == 1 else 25):
        op = ops[int(rng.integers(len(ops)))]
        log.append((seed, scheme, (nx, ny, nz), k, op, transport))
        if op == "step":
            n = int(rng.integers(1, 4))
            for _ in range(n):
                w.collision_MRT(); w.macrovar()
            sim.run_device(n)
        elif op == "download":
            sim.download_f(out)
            ref = w.get_f()[sl]
            assert np.array_equal(out, ref) if exact else np.max(np.abs(out - ref)) <= rtol * np.max(np.abs(w.get_f()))
        elif op == "macro":
            sim.device_macrovar()
            for name in ("rho", "ux", "uy", "uz"):
                a, b = getattr(sim, name), w.get(name)[sl]
                assert np.array_equal(a, b) if exact else np.max(np.abs(a - b)) <= rtol * max(np.max(np.abs(w.get_f())), 1e-300), name
        elif op == "probe":
            ix, iy, iz = (int(rng.integers(1, n + 1)) for n in (nx, ny, nz))
            if sim.globalz < iz <= sim.globalz + sim.lz:
                pr = sim.probe(ix, iy, iz - sim.globalz)
                ref = np.array([w.get(name)[iz - 1, iy - 1, ix - 1] for name in ("rho", "ux", "uy", "uz")])
                assert np.array_equal(pr, ref) if exact else np.max(np.abs(pr - ref)) <= rtol
        elif op == "sums":
            got = sim.profiles()
            ref, _ = orc.plane_sums(w)
            scale = np.max(np.abs(ref), axis=1, keepdims=True) + 1e-30
            assert np.max(np.abs(got - ref) / np.maximum(scale, np.max(np.abs(ref[1])) * 1e-3)) < 1e-11
        elif op == "reupload":
            f = w.get_f() + 1e-5 * rng.normal(size=shp + (19,))
            w.set_f(f); w.macrovar()
            sim.upload_f(np.ascontiguousarray(f[sl]))
            exact = not fast
        elif op == "force":
            F = [float(t) for t in 1e-6 * rng.normal(size=3)]
            for name, v in zip(("fx", "fy", "fz"), F):
                w.set(name, np.full(shp, v))
            sim.set_force_uniform(*F)                      # also drops a force field
            w.macrovar()
        elif op == "field":
            F = [1e-6 * rng.normal(size=shp) for _ in range(3)]
            for name, a in zip(("fx", "fy", "fz"), F):
                w.set(name, a)
            sim.set_force_field(*[np.ascontiguousarray(a[sl]) for a in F])
            w.macrovar()
        elif op == "external":
            macro = [1e-4 * rng.normal(size=shp)] + [0.01 * rng.normal(size=shp) for _ in range(3)]
            for name, a in zip(("rho", "ux", "uy", "uz"), macro):
                w.set(name, a)
            sim.set_macro(*[np.ascontiguousarray(a[sl]) for a in macro])
            w.collision_MRT(); w.macrovar()
            sim.collide_stream(capi.MACRO_EXTERNAL)
        elif op == "prerelax":
            # rhoupdat; collision with frozen u (main.f90:74-76): the arrays must hold what macrovar left
            sim.device_macrovar(download=False)
            w.rhoupdat(); w.collision_MRT(); w.macrovar()
            sim.collide_stream(capi.MACRO_PRERELAX)
        elif op == "avedensity":
            sim.device_macrovar(download=False)
            mean_ref, n_ref = w.avedensity()
            m, n = C.c_double(0), C.c_int64(0)
            capi.check(sim.L.d3q19_avedensity(sim.h, C.byref(m), C.byref(n)))
            assert n.value == n_ref and abs(m.value - mean_ref) <= rtol * max(np.max(np.abs(w.get_f())), 1e-300)
            w.collision_MRT(); w.macrovar()
            sim.collide_stream()
            exact = False
    sim.download_f(out)
    ref = w.get_f()[sl]
    assert np.array_equal(out, ref) if exact else np.max(np.abs(out - ref)) <= 10 * rtol * np.max(np.abs(w.get_f()))
    sim.close(); w.close()


def run_shim_sequence(seed, scheme, log):
    """the driver-facing entry points (what collision_b200.f90 binds) in random but legal order: segments of the time
    loop with random output cadences, a host-side rewrite of f (loadcntdflow) at a random point, pre-relaxation
    pairs, sync_f_to_host (savecntdflow), avedensity steps -- the host arrays must be current exactly when the intact
    driver would read them"""
    rng = np.random.default_rng(1000 + seed)
    nx, ny, nz = int(rng.integers(6, 30)), int(rng.integers(1, 5)), int(rng.integers(1, 5))
    U = dict(ustar=0.0025, ystar=0.0036 / 0.0025, force_in_y=2.0 * 0.0025 * 0.0025 / nx)
    ndiag, nflowout = int(rng.integers(2, 7)), int(rng.integers(2, 7))
    w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True, **U)
    sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, scheme=scheme, math_mode=capi.MATH_STRICT, ndiag=ndiag, nflowout=nflowout, **U)
    sim.f[...] = w.get_f()
    for k in ("rho", "ux", "uy", "uz"):
        getattr(sim, k)[...] = w.get(k)
    sim.host_f_changed()
    sim.FORCING()
    names = ("rho", "ux", "uy", "uz")

    def host_macro_current(what):
        for k in names:
            assert np.array_equal(getattr(sim, k), w.get(k)), (what, k)

    for seg in range(6):
        op = ["loop", "loop", "rewrite", "prerelax", "save", "checkpoint"][int(rng.integers(6))]
        log.append((seed, scheme, (nx, ny, nz), seg, "shim:" + op))
        if op == "loop":
            nsteps = int(rng.integers(1, 14))
            sim.v.nsteps = nsteps
            sim.v.istep0 = int(rng.integers(0, 50))
            sim.set_schedule()
            # main.f90:132-136 before every (re)start of the loop
            sim.FORCING(); w.macrovar()
            sim.istep = sim.v.istep0
            sim.macrovar()
            host_macro_current("initial macrovar")
            for step in range(sim.v.istep0 + 1, sim.v.istep0 + nsteps + 1):
                sim.istep = step
                w.collision_MRT(); w.macrovar()
                sim.collision_MRT(); sim.macrovar()
                if step % ndiag == 0 or step % nflowout == 0 or step == sim.v.istep0 + nsteps:
                    host_macro_current("step %d" % step)
        elif op == "rewrite":
            f = w.get_f() + 1e-5 * rng.normal(size=(nz, ny, nx, 19))
            w.set_f(f)
            sim.f[...] = f
            sim.host_f_changed()
        elif op == "prerelax":
            # the driver's arrays hold what the last macrovar left; u stays frozen (main.f90:70-90)
            w.macrovar()
            sim.istep = 0
            sim.macrovar()
            for it in range(int(rng.integers(1, 5))):
                w.rhoupdat(); w.collision_MRT()
                sim.rhoupdat(); sim.collision_MRT()
                assert np.array_equal(sim.rho, w.get("rho")), "rho after rhoupdat"
        elif op == "checkpoint":
            # savecntdflow (saveload.f90:196-231), then a NEW run on the other storage scheme restarts from the file
            import tempfile
            with tempfile.TemporaryDirectory() as d:
                pkg.saveload.savecntdflow(sim, d, istat=3, imovie=1)
                istep = sim.v.istep0 + sim.v.nsteps
                sim.close()
                scheme = capi.SCHEME_AB if scheme == capi.SCHEME_AA else capi.SCHEME_AA
                sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, scheme=scheme, math_mode=capi.MATH_STRICT, ndiag=ndiag,
                                      nflowout=nflowout, **U)
                assert pkg.saveload.loadcntdflow(sim, d, istep) == (istep, 3, 1)
            sim.FORCING()
            assert np.array_equal(sim.f, w.get_f()), "checkpoint payload"
        else:
            got = sim.sync_f_to_host()
            assert np.array_equal(got, w.get_f()), "sync_f_to_host"
    assert np.array_equal(sim.sync_f_to_host(), w.get_f())
    sim.close(); w.close()


def run_particle_sequence(seed, scheme, log, rank=0, world=1, comm=None):
    """the particle entry points in random order: fixed and moving particle steps, macrovar with the solid branch,
    avedensity over the fluid nodes, diag and the masked plane sums, link list and mask read-backs -- against
    oracle/particles_oracle.c + the fluid oracle (parity unpinned against the reference: partlib.f90 is absent)"""
    from oracle import particles as P
    rng = np.random.default_rng(2000 + seed)
    nx, ny, nz, rad = 20, 16, 18 if world == 1 else 7 * world, 3.3
    U = dict(ustar=0.0025, ystar=0.0036 / 0.0025, force_in_y=2.0 * 0.0025 * 0.0025 / nx)
    pos = [[9.7, 1.2 + 3 * rng.random(), nz - 1.1], [5.1 + 2 * rng.random(), 10.0, 7.2]]      # both cut by slab faces
    vel = [list(0.02 * (rng.random(3) - 0.5)) for _ in range(2)]
    omg = [list(2e-3 * (rng.random(3) - 0.5)) for _ in range(2)]
    w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True, ipart=1, **U)
    sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, scheme=scheme, math_mode=capi.MATH_FAST, ipart=True, rank=rank,
                          nranks=world, device=0, nccl_id=comm.new_id(rank) if world > 1 else None, **U)
    sl = slice(sim.globalz, sim.globalz + sim.lz)
    sim.FORCING()
    sim.upload_f(np.ascontiguousarray(w.get_f()[sl]))
    pt = P.Particles(nx, ny, nz, rad, pos, vel, omg)
    sim.particles_init(pos, rad, vel, omg)
    pt.build_mask(); pt.build_links()

    def oracle_mask():
        w.set_solid(np.where(pt.own > 0, 1, -1).astype(np.int32), pt.own)
        w.set_particles(pt.ypglb, pt.wp, pt.omgp)
    oracle_mask()
    w.macrovar()
    out = np.empty((sim.lz, ny, nx, 19))
    tol = 1e-11

    def fields_agree(what, f):
        nonlocal tol
        fluid = pt.own[sl] < 0
        sim.download_f(out)
        err = np.max(np.abs(out[fluid] - f[sl][fluid])) / np.max(np.abs(f[pt.own < 0]))
        assert err < tol, (what, err)

    for seg in range(7):
        op = ["fixed", "moving", "macro", "avedensity", "monitors", "links"][int(rng.integers(6))]
        log.append((seed, scheme, (nx, ny, nz), seg, "particles:" + op))
        if op == "fixed":
            for _ in range(int(rng.integers(1, 4))):
                w.collision_MRT()
                f = w.get_f(); pt.ibb(f); w.set_f(f); w.macrovar()
                sim.particle_step(move=False)
            fields_agree("fixed", f)
            g = sim.get_particles()
            assert np.max(np.abs(g["fHIp"] - pt.fHIp)) < 1e-8 * np.max(np.abs(pt.fHIp))
        elif op == "moving":
            for _ in range(int(rng.integers(1, 4))):
                w.collision_MRT()
                f = w.get_f(); pt.ibb(f); pt.lubforce(); pt.move(); pt.build_mask(); pt.build_links(); pt.refill(f)
                w.set_f(f); oracle_mask(); w.macrovar()
                sim.particle_step(move=True)
            tol = 1e-8                                   # refills extrapolate: differences of 1e-12 grow a little
            g = sim.get_particles()
            assert np.max(np.abs(g["ypglb"] - pt.ypglb)) < 1e-10
            assert np.array_equal(sim.get_mask(), pt.own[sl])
            fields_agree("moving", f)
        elif op == "macro":
            sim.device_macrovar()
            solid = pt.own[sl] > 0
            for name in ("rho", "ux", "uy", "uz"):
                a, b = getattr(sim, name), w.get(name)[sl]
                assert np.max(np.abs(a[~solid] - b[~solid])) <= tol * max(np.max(np.abs(b)), 1e-30), name
                if solid.any():                           # (a thin slab may hold no solid node)
                    assert np.max(np.abs(a[solid] - b[solid])) <= 1e-9 * max(np.max(np.abs(b)), 1e-30), name   # rigid-body velocity
        elif op == "avedensity":
            sim.device_macrovar(download=False)
            mean_ref, n_ref = w.avedensity()
            m, n = C.c_double(0), C.c_int64(0)
            capi.check(sim.L.d3q19_avedensity(sim.h, C.byref(m), C.byref(n)))
            assert n.value == n_ref == int((pt.own < 0).sum())
            assert abs(m.value - mean_ref) <= 1e-9 * max(np.max(np.abs(w.get_f())), 1e-300)
            w.macrovar()                                 # rho recomputed: the shift lives in the arrays only
            sim.L.d3q19_macrovar(sim.h)
        elif op == "monitors":
            d, ref = sim.diag(), orc.diag_line(w, p.ustar, solid=pt.own > 0)
            assert d["nfluid"] == ref["nfluid"] and abs(d["volf"] - ref["volf"]) < 1e-15
            assert abs(d["vmax"] - ref["vmax"]) <= 1e-7 * ref["vmax"]
            p2 = sim.profiles2()
            sums, cnt = orc.plane_sums(w, solid=pt.own > 0)
            assert np.array_equal(p2[11], cnt)
            assert np.max(np.abs(p2[1] - sums[1])) <= 1e-7 * np.max(np.abs(sums[1]))
        else:
            n = sim.beads_links()
            gl = sim.get_links()
            mine = (pt.links["z"] > sim.globalz) & (pt.links["z"] <= sim.globalz + sim.lz)      # the links whose fluid node I own
            assert n == int(mine.sum())
            want = P.canon(pt.links, mine)
            for key in ("x", "y", "z", "ip", "part"):
                assert np.array_equal(gl[key], want[key]), key
            assert np.array_equal(sim.get_mask(), pt.own[sl])
    sim.close(); w.close()


def main():
    if "hostsim" not in os.path.basename(capi.LIB_PATH):
        print("this script drives the host-sim build only (set D3Q19_LIB)")
        return 2
    nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    log = []
    try:
        for seed in range(nseeds):
            for scheme in (capi.SCHEME_AA, capi.SCHEME_AB):
                run_sequence(seed, scheme, log)
                run_shim_sequence(seed, scheme, log)
                if seed < int(os.environ.get("HOSTSIM_PARTICLE_SEEDS", "2")):
                    run_particle_sequence(seed, scheme, log)
    except Exception:
        import traceback
        traceback.print_exc()
        print("last calls:", log[-8:])
        print("HOSTSIM_RANDOM_FAILED")
        return 1
    print("HOSTSIM_RANDOM_OK %d sequences" % (4 * nseeds))
    return 0


if __name__ == "__main__":
    sys.exit(main())
