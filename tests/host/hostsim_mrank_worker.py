"""TEST INFRASTRUCTURE ONLY.  The multi-GPU parity worker (tests/mgpu_worker.py) for the host-sim build of the library
(tests/host/make_hostsim.py): every rank is a THREAD of this process, NCCL is tests/host/fake/fake_nccl.h, peer memory
is the process's own address space.  Run as a subprocess with D3Q19_LIB pointing at the host-sim library:

    D3Q19_LIB=tests/host/_gen/libd3q19b200_hostsim.so python tests/host/hostsim_mrank_worker.py <world> [section ...]

What it shows: the REAL orchestration of csrc/d3q19_api.cu (slab geometry, face exchange, send-back after odd steps,
peer-memory connect + flag protocol, copy-engine put transport, reductions, particle link partition, force
all-reduce, refill source exchange) computes, on 2-4 slabs, bit for bit what the single-domain oracle computes.
What it cannot show: ordering between CUDA streams (the fake device is synchronous; see test_halo_schedule_model.py).
"""
import os
import sys
import threading
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle import particles as P  # noqa: E402

pkg = entry.load_package()
capi = pkg.capi


class Comm:
    """what torch.distributed does for the real worker"""

    def __init__(self, world):
        self.world, self.bar, self.slots, self.lock = world, threading.Barrier(world), [None] * world, threading.Lock()
        self.failed = []

    def allgather(self, rank, item):
        self.bar.wait()
        self.slots[rank] = item
        self.bar.wait()
        out = list(self.slots)
        self.bar.wait()
        return out

    def new_id(self, rank):
        return self.allgather(rank, bytes(capi.nccl_unique_id()) if rank == 0 else None)[0]

    def fail(self, rank, what):
        with self.lock:
            self.failed.append("rank %d: %s" % (rank, what))
            print("rank %d: %s" % (rank, what), flush=True)       # at once: the other ranks may hang in a collective now


def section_fluid(rank, world, comm, chk, ctx):
    # (size, overlap, transport)
    cases = [((24, 6, 4 * world), True, "nccl"), ((33, 5, 3 * world + 1), True, "nccl"), ((16, 4, world), True, "nccl"),
             ((24, 6, 4 * world), False, "nccl"), ((130, 3, 2 * world + 1), True, "nccl"),
             ((24, 6, 4 * world), True, "peer"), ((33, 5, 3 * world + 1), True, "peer"), ((130, 3, 2 * world + 1), True, "peer"),
             ((24, 6, 4 * world), True, "peer-split"), ((33, 5, 3 * world + 1), True, "peer-split"),
             ((24, 6, 4 * world), True, "put"), ((33, 5, 3 * world + 1), True, "put"), ((130, 3, 2 * world + 1), True, "put")]
    if os.environ.get("HOSTSIM_SHORT"):          # the default CPU suite: one uneven case per transport
        cases = [((33, 5, 3 * world + 1), True, h) for h in ("nccl", "peer", "peer-split", "put")] + \
                [((24, 6, 4 * world), False, "nccl")]
    import time
    for scheme in (capi.SCHEME_AA, capi.SCHEME_AB):
        for (nx, ny, nz), overlap, halo in cases:
            t_case = time.perf_counter()
            ctx[0] = "scheme %d case %s overlap %s halo %s" % (scheme, (nx, ny, nz), overlap, halo)
            w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True)
            w.set_f(w.get_f() + 1e-4 * np.random.default_rng(7).normal(size=(nz, ny, nx, 19)))
            sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=0, scheme=scheme,
                                  math_mode=capi.MATH_STRICT, nccl_id=comm.new_id(rank), overlap=overlap,
                                  halo_split_min=3 if halo == "peer-split" else 0)
            z0, z1 = sim.globalz, sim.globalz + sim.lz
            sim.FORCING()
            if halo.startswith("peer") or halo.startswith("put"):
                ok_ = sim.connect_halo(lambda b: comm.allgather(rank, bytes(b)), mode="put" if halo.startswith("put") else "fused")
                chk("peer halo connects", ok_)
            sim.upload_f(np.ascontiguousarray(w.get_f()[z0:z1]))
            w.macrovar()
            out = np.empty((sim.lz, ny, nx, 19))
            for burst in (1, 1, 2, 1, 3):
                for _ in range(burst):
                    w.collision_MRT(); w.macrovar()
                sim.run_device(burst)
                sim.download_f(out)
                # (no early exit: the ranks must stay in lockstep for the collectives that follow)
                chk("populations after a burst of %d" % burst, bool(np.array_equal(out, w.get_f()[z0:z1])))
            sim.device_macrovar()
            for k in ("rho", "ux", "uy", "uz"):
                chk("macrovar " + k, bool(np.array_equal(getattr(sim, k), w.get(k)[z0:z1])))
            if sim.lz >= 2 and ny >= 2:
                for name, got, want in zip(("ox", "oy", "oz"), sim.vortcalc(), w.vortcalc()):
                    chk("vortcalc " + name, bool(np.array_equal(got, want[z0:z1])))
            pr = sim.profiles()
            ref = w.get("uy").sum(axis=(0, 1))
            chk("profiles", bool(np.allclose(pr[1], ref, rtol=1e-12, atol=1e-13 * np.max(np.abs(ref)))))
            d, dref = sim.diag(), orc.diag_line(w, p.ustar)
            chk("diag location", (d["imout"], d["jmout"], d["kmout"]) == (dref["imout"], dref["jmout"], dref["kmout"]))
            chk("diag vmax", d["vmax"] == dref["vmax"] and d["nfluid"] == dref["nfluid"])
            # avedensity across ranks (MPI_ALLREDUCE, collision.f90:500-501), then the shifted collision
            import ctypes as C
            mean_ref, n_ref = w.avedensity()
            m, n = C.c_double(0), C.c_int64(0)
            capi.check(sim.L.d3q19_avedensity(sim.h, C.byref(m), C.byref(n)))
            chk("avedensity count", n.value == n_ref)
            chk("avedensity mean", abs(m.value - mean_ref) <= 1e-12 * float(np.mean(np.abs(w.get("rho")))) + 1e-300)
            w.collision_MRT()
            sim.collide_stream()
            sim.download_f(out)
            err = np.max(np.abs(out - w.get_f()[z0:z1])) / np.max(np.abs(w.get_f()))
            chk("step after avedensity err %g" % err, bool(err < 1e-13))
            sim.close(); w.close()
            if rank == 0 and os.environ.get("HOSTSIM_VERBOSE"):
                print("%-60s %.2f s" % (ctx[0], time.perf_counter() - t_case), flush=True)


def section_prerelax(rank, world, comm, chk, ctx):
    ctx[0] = "device prerelax"
    nx, ny, nz = 32, 8, 4 * world
    tol = 2e-5

    def start():
        w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True)
        w.set_f(w.get_f() + 1e-4 * np.random.default_rng(5).normal(size=(nz, ny, nx, 19)))
        return w, p
    w, p = start()
    it_ref = 0
    while True:
        rhop = w.get("rho").copy()
        w.rhoupdat(); w.collision_MRT()
        err_ref = np.max(np.abs(w.get("rho") - rhop))
        if err_ref <= tol or it_ref > 12:
            break
        it_ref += 1
    w0, _ = start()
    sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=0, math_mode=capi.MATH_STRICT,
                          nccl_id=comm.new_id(rank), rhoepsl=tol)
    z0, z1 = sim.globalz, sim.globalz + sim.lz
    sim.f[...] = w0.get_f()[z0:z1]
    for k in ("rho", "ux", "uy", "uz"):
        getattr(sim, k)[...] = w0.get(k)[z0:z1]
    sim.FORCING()
    it, err = sim.prerelax_device(maxiter=12)
    chk("iterations %d vs %d" % (it, it_ref), it == it_ref)
    chk("rhoerr %r vs %r" % (err, err_ref), err == err_ref)
    chk("f after prerelax", bool(np.array_equal(sim.f, w.get_f()[z0:z1])))
    sim.close()


def section_particles(rank, world, comm, chk, ctx):
    nx, ny, nz, rad = 24, 20, 8 * world, 3.6
    U = dict(ustar=0.0025, ystar=0.0036 / 0.0025, force_in_y=2.0 * 0.0025 * 0.0025 / nx)
    pos = [[11.7, 1.2, 8.0 * world - 0.9], [8.3, 12.0, 8.1], [15.5, 8.4, 4.2]]      # two of them cut by slab faces
    vel = [[0.010, 0.020, -0.010], [0.0, 0.015, 0.0], [-0.005, 0.0, 0.012]]
    omg = [[1e-3, 0.0, 2e-3], [0.0, -1e-3, 0.0], [5e-4, 5e-4, 0.0]]
    for scheme, halo in ((capi.SCHEME_AA, "nccl"), (capi.SCHEME_AB, "nccl"), (capi.SCHEME_AA, "put"), (capi.SCHEME_AB, "put")):
        ctx[0] = "particles scheme %d faces by %s" % (scheme, halo)
        w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True, ipart=1, **U)
        pt = P.Particles(nx, ny, nz, rad, pos, vel, omg)
        pt.build_mask(); pt.build_links()
        w.set_solid(np.where(pt.own > 0, 1, -1).astype(np.int32), pt.own)
        w.set_particles(pt.ypglb, pt.wp, pt.omgp)
        sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=0, scheme=scheme,
                              nccl_id=comm.new_id(rank), ipart=True, **U)
        z0, z1 = sim.globalz, sim.globalz + sim.lz
        sim.FORCING()
        if halo == "put":                              # faces by the copy engines; the refill sources still travel by NCCL
            chk("peer halo connects", sim.connect_halo(lambda b: comm.allgather(rank, bytes(b)), mode="put"))
        sim.upload_f(np.ascontiguousarray(w.get_f()[z0:z1]))
        sim.particles_init(pos, rad, vel, omg)
        nl = sim.beads_links()
        tot = sum(comm.allgather(rank, nl))
        chk("link count %d vs %d" % (tot, len(pt.links["q"])), tot == len(pt.links["q"]))
        chk("mask", bool(np.array_equal(sim.get_mask(), pt.own[z0:z1])))
        gl = sim.get_links()
        mine = P.canon(pt.links, (pt.links["z"] > z0) & (pt.links["z"] <= z1))     # the oracle's links whose fluid node I own
        for key in ("x", "y", "z", "ip", "part", "q"):
            chk("links " + key, bool(np.array_equal(gl[key], mine[key])))
        w.macrovar()
        out = np.empty((sim.lz, ny, nx, 19))
        for step in range(3):
            w.collision_MRT()
            f = w.get_f(); pt.ibb(f); w.set_f(f); w.macrovar()
            sim.particle_step(move=False)
            sim.download_f(out)
            fluid = pt.own[z0:z1] < 0
            err = np.max(np.abs(out[fluid] - f[z0:z1][fluid])) / np.max(np.abs(f))
            chk("ibb step %d err %g" % (step, err), bool(err < 1e-12))
            g = sim.get_particles()
            ferr = np.max(np.abs(g["fHIp"] - pt.fHIp)) / np.max(np.abs(pt.fHIp))
            chk("force step %d err %g" % (step, ferr), bool(ferr < 1e-10))
        for step in range(8):                          # moving, with the refill sources exchanged across the faces
            w.collision_MRT()
            f = w.get_f(); pt.ibb(f); pt.lubforce(); pt.move(); pt.build_mask(); pt.build_links(); pt.refill(f)
            w.set_f(f); w.set_solid(np.where(pt.own > 0, 1, -1).astype(np.int32), pt.own)
            w.set_particles(pt.ypglb, pt.wp, pt.omgp); w.macrovar()
            sim.particle_step(move=True)
            g = sim.get_particles()
            chk("moving positions %d" % step, bool(np.max(np.abs(g["ypglb"] - pt.ypglb)) < 1e-9))
            chk("moving mask %d" % step, bool(np.array_equal(sim.get_mask(), pt.own[z0:z1])))
            ferr = np.max(np.abs(g["fHIp"] - pt.fHIp)) / np.max(np.abs(pt.fHIp))
            chk("moving force %d err %g" % (step, ferr), bool(ferr < 1e-9))
            sim.download_f(out)
            fluid = pt.own[z0:z1] < 0
            err = np.max(np.abs(out[fluid] - f[z0:z1][fluid])) / np.max(np.abs(f))
            chk("moving populations %d err %g" % (step, err), bool(err < 1e-9))
        sim.close(); w.close()


def section_shim(rank, world, comm, chk, ctx):
    """main.f90's own call sequence on every rank, through the entry points the Fortran shim binds: pre-relaxation by
    rhoupdat / collision_MRT (the library predicts the loop exit from the all-reduced max|rho - rhop|, and makes the
    host f current in that iteration for saveinitflow), then the time loop with the diag / output cadence -- the host
    arrays must be current exactly when the driver reads them, on every rank."""
    nx, ny, nz, tol, itmax, nsteps = 23, 6, 3 * world + 1, 2e-5, 15000, 12          # main.f90:85 bounds the loop at 15000
    for scheme in (capi.SCHEME_AA, capi.SCHEME_AB):
        ctx[0] = "shim scheme %d" % scheme

        def start():
            w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True)
            w.set_f(w.get_f() + 1e-4 * np.random.default_rng(5).normal(size=(nz, ny, nx, 19)))
            return w, p
        w, p = start()
        sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=0, scheme=scheme,
                              math_mode=capi.MATH_STRICT, nccl_id=comm.new_id(rank), rhoepsl=tol, ndiag=5, nflowout=4)
        z0, z1 = sim.globalz, sim.globalz + sim.lz
        sim.f[...] = w.get_f()[z0:z1]
        for k in ("rho", "ux", "uy", "uz"):
            getattr(sim, k)[...] = w.get(k)[z0:z1]
        sim.host_f_changed()
        sim.FORCING()
        # main.f90:70-90 on the oracle
        it_ref = 0
        while True:
            rhop = w.get("rho").copy()
            w.rhoupdat(); w.collision_MRT()
            err_ref = float(np.max(np.abs(w.get("rho") - rhop)))
            if err_ref <= tol or it_ref > itmax:
                break
            it_ref += 1
        it, err = sim.prerelax(allreduce_max=lambda x: max(comm.allgather(rank, x)), maxiter=itmax)
        chk("prerelax iterations %d vs %d" % (it, it_ref), it == it_ref)
        chk("prerelax error %r vs %r" % (err, err_ref), err == err_ref)
        chk("host f current for saveinitflow", bool(np.array_equal(sim.f, w.get_f()[z0:z1])))
        chk("host rho current", bool(np.array_equal(sim.rho, w.get("rho")[z0:z1])))
        # main.f90:102-136, then the loop :142-208
        w.macrovar()
        sim.macrovar()
        sim.istep = 0
        sim.FORCING()
        sim.macrovar()
        for k in ("rho", "ux", "uy", "uz"):
            chk("initial macrovar " + k, bool(np.array_equal(getattr(sim, k), w.get(k)[z0:z1])))
        seen = []

        def on_step(s_):
            if s_.istep % s_.v.ndiag == 0 or s_.istep % s_.v.nflowout == 0 or s_.istep == nsteps:
                seen.append((s_.istep, [getattr(s_, k).copy() for k in ("rho", "ux", "uy", "uz")]))
        ref = {}
        sim.v.istep0 = 0
        for step in range(1, nsteps + 1):
            w.collision_MRT(); w.macrovar()
            if step % 5 == 0 or step % 4 == 0 or step == nsteps:
                ref[step] = [w.get(k)[z0:z1].copy() for k in ("rho", "ux", "uy", "uz")]
        sim.run(nsteps, on_step=on_step)
        chk("output steps %r" % [t for t, _ in seen], [t for t, _ in seen] == sorted(ref))
        for t, arrs in seen:
            chk("host rho,u current on step %d" % t, all(bool(np.array_equal(a, b)) for a, b in zip(arrs, ref[t])))
        chk("f after the loop", bool(np.array_equal(sim.sync_f_to_host(), w.get_f()[z0:z1])))
        sim.close(); w.close()


def section_restart(rank, world, comm, chk, ctx):
    """checkpoint / restart over slabs (saveload.f90:196-231, :296-332: one file per rank), into the other storage scheme"""
    import tempfile
    saveload = pkg.saveload
    nx, ny, nz = 24, 6, 3 * world + 1
    tmp = comm.allgather(rank, tempfile.mkdtemp(prefix="d3q19_restart_") if rank == 0 else None)[0]
    for first, second in ((capi.SCHEME_AA, capi.SCHEME_AB), (capi.SCHEME_AB, capi.SCHEME_AA)):
        ctx[0] = "restart %d -> %d" % (first, second)
        w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True)
        a = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=0, scheme=first,
                            math_mode=capi.MATH_STRICT, nccl_id=comm.new_id(rank))
        z0, z1 = a.globalz, a.globalz + a.lz
        a.f[...] = w.get_f()[z0:z1]
        a.host_f_changed(); a.FORCING(); a.macrovar()
        a.run(9)
        saveload.savecntdflow(a, tmp)
        a.close()
        comm.bar.wait()
        b = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=0, scheme=second,
                            math_mode=capi.MATH_STRICT, nccl_id=comm.new_id(rank))
        b.FORCING()
        chk("istep0", saveload.loadcntdflow(b, tmp, 9)[0] == 9)
        b.macrovar()
        b.run(11)
        w.macrovar()
        for _ in range(20):
            w.collision_MRT(); w.macrovar()
        chk("20 steps with a restart after 9", bool(np.array_equal(b.sync_f_to_host(), w.get_f()[z0:z1])))
        b.close(); w.close()
        comm.bar.wait()


def section_benchparity(rank, world, comm, chk, ctx):
    """bench.py's own pre-timing parity check (the reference's golden vector on this job's slabs, every transport; the
    moving-particle case against one domain), with the ranks as threads"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for halo in ("nccl", "peer", "put"):
        ctx[0] = "bench parity_check, halo " + halo

        def connect(sim, particles=False):
            if halo != "nccl":
                assert sim.connect_halo(lambda b: comm.allgather(rank, bytes(b)), mode="put" if (halo == "put" or particles) else "fused")

        res = bench.parity_check(pkg, rank, world, 0, lambda: comm.new_id(rank), connect,
                                 lambda ok: all(comm.allgather(rank, bool(ok))), lambda obj: comm.allgather(rank, obj)[0],
                                 particles=halo != "peer")
        chk("parity_check %r" % (res,), res["bit_exact"] is True and len(res["schemes"]) == 2)
        if halo != "peer":
            chk("particle leg present", res["particles"]["ok"] is True)


def section_random(rank, world, comm, chk, ctx):
    """random call sequences (tests/host/hostsim_random_calls.py) followed by every rank, one transport per sequence"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("hostsim_random_calls", os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                                                     "hostsim_random_calls.py"))
    rc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rc)
    transports = ["nccl", "peer", "peer-split", "put"]
    nseeds = int(os.environ.get("HOSTSIM_RANDOM_SEEDS", "7"))
    for seed in range(nseeds):
        for scheme in (capi.SCHEME_AA, capi.SCHEME_AB):
            tr = transports[(seed + scheme) % len(transports)]
            ctx[0] = "random seed %d scheme %d transport %s" % (seed, scheme, tr)
            log = []
            try:
                rc.run_sequence(500 + seed, scheme, log, rank=rank, world=world, comm=comm, transport=tr)
            except Exception:
                chk("random sequence, last calls %r\n%s" % (log[-6:], traceback.format_exc()), False)
                raise                      # the ranks have left lockstep: stop the run
    for seed in range(int(os.environ.get("HOSTSIM_PARTICLE_SEEDS", "1"))):
        for scheme in (capi.SCHEME_AA, capi.SCHEME_AB):
            ctx[0] = "random particle sequence seed %d scheme %d" % (seed, scheme)
            log = []
            try:
                rc.run_particle_sequence(seed, scheme, log, rank=rank, world=world, comm=comm)
            except Exception:
                chk("random particle sequence, last calls %r\n%s" % (log[-6:], traceback.format_exc()), False)
                raise


SECTIONS = {"random": section_random, "benchparity": section_benchparity, "restart": section_restart, "fluid": section_fluid, "prerelax": section_prerelax, "particles": section_particles, "shim": section_shim}


def rank_main(rank, world, comm, sections):
    ctx = [""]

    def chk(name, cond):
        if not cond:
            comm.fail(rank, "CHECK FAILED %s [%s]" % (name, ctx[0]))
        return cond
    try:
        for s in sections:
            SECTIONS[s](rank, world, comm, chk, ctx)
    except Exception:
        comm.fail(rank, "EXCEPTION [%s]\n%s" % (ctx[0], traceback.format_exc()))
        # the other ranks are (or will be) waiting for this one in a barrier or a collective: end the whole run here
        print("HOSTSIM_MRANK_FAILED", flush=True)
        os._exit(1)


def main():
    world = int(sys.argv[1])
    sections = sys.argv[2:] or list(SECTIONS)
    if "hostsim" not in os.path.basename(capi.LIB_PATH):
        print("this worker drives the host-sim build only (set D3Q19_LIB)")
        return 2
    comm = Comm(world)
    threads = [threading.Thread(target=rank_main, args=(r, world, comm, sections)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for line in comm.failed[:20]:
        print(line)
    print("HOSTSIM_MRANK_OK" if not comm.failed else "HOSTSIM_MRANK_FAILED")
    return 0 if not comm.failed else 1


if __name__ == "__main__":
    sys.exit(main())
