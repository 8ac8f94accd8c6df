"""Multi-GPU parity worker (launched by tests/test_gpu_multi.py under torch.distributed.run).

Each rank owns one z-slab on its own GPU.  The global result must be BIT-IDENTICAL to the
single-domain oracle (streaming is a copy, collision is node-local), for
both storage schemes, with and without the boundary/interior overlap, for even and uneven
slabs; scalar reductions (avedensity, pre-relaxation error, profiles) must agree over ranks.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
from oracle import oracle as orc  # noqa: E402

pkg = entry.load_package()
capi = pkg.capi


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    torch.cuda.set_device(local)
    ok = True
    ctx = [""]

    def chk(name, cond):
        nonlocal ok
        if not cond:
            print("rank %d: CHECK FAILED %s [%s]" % (rank, name, ctx[0]), flush=True)
            ok = False
        return cond

    def new_id():
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.tensor(list(capi.nccl_unique_id()), dtype=torch.uint8)
        dist.broadcast(idt, src=0)
        return bytes(idt.tolist())

    def allgather_bytes(b):
        t = torch.tensor(list(b), dtype=torch.uint8)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [bytes(o.tolist()) for o in out]

    # (size, overlap, halo): halo = "peer" stores the faces straight into the neighbour GPU's memory
    cases = [((24, 6, 4 * world), True, "nccl"), ((33, 5, 3 * world + 1), True, "nccl"), ((16, 4, world), True, "nccl"),
             ((24, 6, 4 * world), False, "nccl"), ((40, 3, 2 * world), True, "nccl"),
             ((24, 6, 4 * world), True, "peer"), ((33, 5, 3 * world + 1), True, "peer"), ((40, 3, 2 * world), True, "peer"),
             ((130, 7, 2 * world + 1), True, "peer"),
             # thick-slab mode of the peer halo: boundary planes and interior as two launches
             ((24, 6, 4 * world), True, "peer-split"), ((33, 5, 3 * world + 1), True, "peer-split"),
             # third transport: plain step kernels, the copy engines move the faces into the neighbours' arrays
             ((24, 6, 4 * world), True, "put"), ((33, 5, 3 * world + 1), True, "put"), ((40, 3, 2 * world), True, "put"),
             ((130, 7, 2 * world + 1), True, "put")]
    only = os.environ.get("MGPU_ONLY", "")          # e.g. "peer": just the peer-memory halo cases (short runs on many GPUs)
    if only:
        cases = [c for c in cases if c[2].startswith(only) or (only == "peer" and c[2].startswith("put"))]
    for scheme in (capi.SCHEME_AA, capi.SCHEME_AB):
        for (nx, ny, nz), overlap, halo in cases:
            ctx[0] = "scheme %d case %s overlap %s halo %s" % (scheme, (nx, ny, nz), overlap, halo)
            w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True)
            rng = np.random.default_rng(7)
            w.set_f(w.get_f() + 1e-4 * rng.normal(size=(nz, ny, nx, 19)))
            sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=local, scheme=scheme,
                                  math_mode=capi.MATH_STRICT, nccl_id=new_id(), overlap=overlap,
                                  halo_split_min=3 if halo == "peer-split" else 0)
            z0, z1 = sim.globalz, sim.globalz + sim.lz
            sim.FORCING()
            if halo.startswith("peer") or halo.startswith("put"):
                connected = sim.connect_halo(allgather_bytes, mode="put" if halo.startswith("put") else "fused")
                if not connected:
                    raise RuntimeError("peer-memory halo unavailable between the GPUs of this box: " + ctx[0])
            sim.upload_f(np.ascontiguousarray(w.get_f()[z0:z1]))
            w.macrovar()
            out = np.empty((sim.lz, ny, nx, 19))
            for step in range(1, 8):
                w.collision_MRT()
                w.macrovar()
                sim.collide_stream()
                sim.download_f(out)
                if not np.array_equal(out, w.get_f()[z0:z1]):
                    print("rank %d: MISMATCH scheme %d case %s step %d" % (rank, scheme, (nx, ny, nz, overlap), step))
                    ok = False
                    break
                if step in (3, 4):
                    sim.device_macrovar()
                    for k in ("rho", "ux", "uy", "uz"):
                        chk("macrovar " + k, bool(np.array_equal(getattr(sim, k), w.get(k)[z0:z1])))
                    if sim.lz >= 2 and ny >= 2:              # vortcalc + exchng8's z phase over NCCL: bit-exact
                        for name, got, want in zip(("ox", "oy", "oz"), sim.vortcalc(), w.vortcalc()):
                            chk("vortcalc " + name, bool(np.array_equal(got, want[z0:z1])))
                    pr = sim.profiles()
                    ref = w.get("uy").sum(axis=(0, 1))
                    chk("profiles", bool(np.allclose(pr[1], ref, rtol=1e-12, atol=1e-13 * np.max(np.abs(ref)))))
            # avedensity across ranks (MPI_ALLREDUCE, collision.f90:500-501)
            sim.device_macrovar()
            rho_scale = float(np.mean(np.abs(w.get("rho"))))
            mean_ref, n_ref = w.avedensity()
            import ctypes as C
            m, n = C.c_double(0), C.c_int64(0)
            capi.check(sim.L.d3q19_avedensity(sim.h, C.byref(m), C.byref(n)))
            chk("avedensity count", n.value == n_ref)
            # the sum is order-dependent: tolerance relative to mean |rho|
            chk("avedensity mean %r vs %r" % (m.value, mean_ref), abs(m.value - mean_ref) <= 1e-12 * rho_scale)
            w.collision_MRT()
            sim.collide_stream()
            sim.download_f(out)
            err = np.max(np.abs(out - w.get_f()[z0:z1])) / np.max(np.abs(w.get_f()))
            chk("step after avedensity err %g" % err, bool(err < 1e-13))
            sim.close()
            w.close()
            if not ok:
                break
        if not ok:
            break

    # device pre-relaxation across ranks (main.f90:70-90 with MPI_ALLREDUCE MAX)
    if ok and not only:
        nx, ny, nz = 32, 8, 4 * world
        w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True)
        w0f, w0 = w.get_f().copy(), {k: w.get(k).copy() for k in ("rho", "ux", "uy", "uz")}
        it_ref = 0
        while True:
            rhop = w.get("rho").copy()
            w.rhoupdat(); w.collision_MRT()
            err_ref = np.max(np.abs(w.get("rho") - rhop))
            if err_ref <= 1e-5 or it_ref > 100:
                break
            it_ref += 1
        sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=local,
                              math_mode=capi.MATH_STRICT, nccl_id=new_id())
        z0, z1 = sim.globalz, sim.globalz + sim.lz
        sim.f[...] = w0f[z0:z1]
        for k in ("rho", "ux", "uy", "uz"):
            getattr(sim, k)[...] = w0[k][z0:z1]
        sim.FORCING()
        it, err = sim.prerelax_device(maxiter=100)
        ctx[0] = "device prerelax"
        chk("iterations %d vs %d" % (it, it_ref), it == it_ref)
        chk("rhoerr %r vs %r" % (err, err_ref), err == err_ref)
        chk("f after prerelax", bool(np.array_equal(sim.f, w.get_f()[z0:z1])))
        sim.close()

    # particle path across slab faces: links partition exactly, IBB + forces as on one domain; moving particles: the
    # refill takes its source nodes across a slab face from the neighbour's planes (exchange in d3q19_beads_filling), so
    # populations, positions and forces stay at the single-domain tolerance (seen on 2 GPUs: profiles/r02b_pytest_gpu_2gpu.log)
    tight = True
    if ok and (not only or only == "particles"):
        from oracle import particles as P
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
            sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=local, scheme=scheme,
                                  nccl_id=new_id(), ipart=True, **U)
            z0, z1 = sim.globalz, sim.globalz + sim.lz
            sim.FORCING()
            if halo == "put" and not sim.connect_halo(allgather_bytes, mode="put"):
                raise RuntimeError("peer-memory halo unavailable between the GPUs of this box: " + ctx[0])
            sim.upload_f(np.ascontiguousarray(w.get_f()[z0:z1]))
            sim.particles_init(pos, rad, vel, omg)
            nl = sim.beads_links()
            tot = torch.tensor([nl]); dist.all_reduce(tot)
            chk("link count %d vs %d" % (int(tot.item()), len(pt.links["q"])), int(tot.item()) == len(pt.links["q"]))
            chk("mask", bool(np.array_equal(sim.get_mask(), pt.own[z0:z1])))
            gl = sim.get_links()
            mine = P.canon(pt.links, (pt.links["z"] > z0) & (pt.links["z"] <= z1))     # the oracle's links whose fluid node I own
            for key in ("x", "y", "z", "ip", "part"):
                chk("links " + key, bool(np.array_equal(gl[key], mine[key])))
            chk("links q", bool(np.array_equal(gl["q"], mine["q"])))
            w.macrovar()
            out = np.empty((sim.lz, ny, nx, 19))
            for step in range(4):
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
            for step in range(10 if tight else 3):                  # moving: positions, mask and forces stay together
                w.collision_MRT()
                f = w.get_f(); pt.ibb(f); pt.lubforce(); pt.move(); pt.build_mask(); pt.build_links(); pt.refill(f)
                w.set_f(f); w.set_solid(np.where(pt.own > 0, 1, -1).astype(np.int32), pt.own)
                w.set_particles(pt.ypglb, pt.wp, pt.omgp); w.macrovar()
                sim.particle_step(move=True)
                g = sim.get_particles()
                chk("moving positions %d" % step, bool(np.max(np.abs(g["ypglb"] - pt.ypglb)) < 1e-9))
                chk("moving mask %d" % step, bool(np.array_equal(sim.get_mask(), pt.own[z0:z1])))
                ferr = np.max(np.abs(g["fHIp"] - pt.fHIp)) / np.max(np.abs(pt.fHIp))
                chk("moving force %d err %g" % (step, ferr), bool(ferr < (1e-9 if tight else 1e-6)))
                if tight:
                    sim.download_f(out)
                    fluid = pt.own[z0:z1] < 0
                    err = np.max(np.abs(out[fluid] - f[z0:z1][fluid])) / np.max(np.abs(f))
                    chk("moving populations %d err %g" % (step, err), bool(err < 1e-9))
            sim.close(); w.close()

    # checkpoint / restart over slabs (saveload.f90:196-231, :296-332: one file per rank): 9 steps, savecntdflow, NEW handles in
    # the other storage scheme, loadcntdflow, 11 more steps == 20 uninterrupted steps of the oracle, bit for bit
    if ok and (not only or only == "restart"):
        import tempfile
        saveload = pkg.saveload
        nx, ny, nz = 24, 6, 3 * world + 1
        tmp = [tempfile.mkdtemp(prefix="d3q19_restart_") if rank == 0 else None]
        dist.broadcast_object_list(tmp, src=0)
        for first, second in ((capi.SCHEME_AA, capi.SCHEME_AB), (capi.SCHEME_AB, capi.SCHEME_AA)):
            ctx[0] = "restart %d -> %d" % (first, second)
            w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True)
            a = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=local, scheme=first,
                                math_mode=capi.MATH_STRICT, nccl_id=new_id())
            z0, z1 = a.globalz, a.globalz + a.lz
            a.f[...] = w.get_f()[z0:z1]
            a.host_f_changed(); a.FORCING(); a.macrovar()
            a.run(9)
            saveload.savecntdflow(a, tmp[0])
            a.close()
            dist.barrier()
            b = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=local, scheme=second,
                                math_mode=capi.MATH_STRICT, nccl_id=new_id())
            b.FORCING()
            chk("istep0", saveload.loadcntdflow(b, tmp[0], 9)[0] == 9)
            b.macrovar()
            b.run(11)
            w.macrovar()
            for _ in range(20):
                w.collision_MRT(); w.macrovar()
            chk("20 steps with a restart after 9", bool(np.array_equal(b.sync_f_to_host(), w.get_f()[z0:z1])))
            b.close(); w.close()
            dist.barrier()

    # halo watchdog: the last rank never steps; a neighbour waiting for its flag must give up after the
    # timeout and d3q19_sync must say so (kernels.cuh halo_spin) -- a dead rank may not hang the others' GPUs
    if ok and only != "particles":
        ctx[0] = "halo watchdog"
        nx, ny, nz = 24, 6, 4 * world
        sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=local, scheme=capi.SCHEME_AB,
                              nccl_id=new_id(), allocate_host=False, halo_timeout_s=2)
        sim.FORCING()
        if not sim.connect_halo(allgather_bytes):
            raise RuntimeError("peer-memory halo unavailable")
        sim.init_channel_device(A9=0.0, noise_amp=1e-4)
        stalled = world - 1
        msg = ""
        if rank != stalled:
            sim.run_device(3)          # step 2 needs the stalled rank's flag of step 1
            try:
                sim.sync()
            except capi.D3Q19Error as exc:
                msg = str(exc)
        dist.barrier()
        is_nb = rank in ((stalled + 1) % world, (stalled - 1) % world) and rank != stalled
        if is_nb:
            chk("neighbour of the stalled rank reports the timeout: %r" % msg, "halo flag" in msg)
        sim.close()

    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_PARITY_OK" if t.item() == 1 else "MGPU_PARITY_FAILED")
    dist.destroy_process_group()
    return 0 if t.item() == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
