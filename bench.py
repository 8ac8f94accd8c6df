#!/usr/bin/env python
"""bench.py -- MLUPS (fp64) of the D3Q19 Channel-Flow time step on N B200s, with the HBM
roofline of the collide-stream kernel and the CPU baseline beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c4|NXxNYxNZ]

A "step" is one collision_MRT (fused collide + force + propagate + wall bounce-back + z-face
exchange) over the whole channel.  At N=1 the workload is BASELINE.json configs[1]:
512x256x256 (nx x ny x nz, x wall-normal), turbulent parameter set, fp64.  For N>1 the
default is weak scaling with one 512x256x256 slab per GPU (global nz = 256 N); `--scaling
strong` keeps the global 512x256x256 (configs[2]).  One JSON line is printed by rank 0.

Other workloads: `--workload c4` = configs[3] (1024x1024x944 per GPU, 150 GB of populations, in place, initialised on the
device), `--particles N` = configs[4] (N moving spheres: links, interpolated bounce-back, force, lubrication, move, refill every
step, avedensity every 100 steps).  With N > 1 the faces travel by the NVLink copy engines (`--halo put`; NCCL send/recv when
peer memory cannot be mapped).

value      device-timed: populations resident in HBM, K steps between CUDA events on the
           stream the kernels run on, max over ranks.
e2e        the same K steps through the reference-facing interface with HOST buffers inside
           the timed region: upload of f(0:18,lx,ly,lz) from pinned host memory, then per step
           collision_MRT + macrovar (the shim's download policy: rho,u come back every
           nflowout/ndiag steps and after the last step) + a `probe` read-back of the centre
           node (32 B D2H) every step.  `upload_s` / `upload_gbs_per_gpu` / `total_s` say where the time goes.
roofline   304 B per node update (19 fp64 in + 19 fp64 out) / mean step-kernel time, against
           MEASURED_PEAKS.json's hbm_gbs; `traffic` = DRAM bytes per launch from the committed ncu capture, reported only
           while the kernel sources are those the capture was taken on (profiles/traffic.json).
parity_check  BEFORE the timed region, on this job's ranks and face transport: a committed golden vector of the reference
           (16 z planes, 9 steps) must come back bit for bit in both storage schemes; with N > 1 a moving-particle case on the
           N slabs must agree with one domain.  A mismatch ends the run non-zero with nothing timed.
clocks     SM clock, power and throttle reasons sampled in process through NVML every 5 ms during the timed region.
cpu_baseline / --impl reference: the reference's own hot path on all host cores, one thread per emulated MPI rank, on a bounded
           sample of the same workload: oracle/_ref/libref_fast.so (the Fortran machine-translated to C, kind "reference"), else
           the hand restatement (kind "port").  (The Fortran itself cannot be built here: no Fortran compiler, no MPI.)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_NODE = 304.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000,
                    help="timed steps (default 1000 = 1.5 s on one B200: long enough for several clock samples, and the "
                         "reference's own nsteps, para.f90:43, so that e2e sees the driver's output cadence)")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--scheme", default="auto", choices=["auto", "aa", "ab"])
    ap.add_argument("--math", default="fast", choices=["fast", "strict"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--no-parity", action="store_true",
                    help="skip the bit-exact check of the reference's golden vector on this job's ranks before the timed region")
    ap.add_argument("--halo", default="put", choices=["peer", "put", "nccl"],
                    help="z-face exchange: put (default) = the copy engines move the faces into the neighbour GPU's memory over "
                         "NVLink next to the interior launch, no SM involved (fastest in every configuration measured, "
                         "profiles/r02e_two_gpus.md; falls back to nccl when peer memory cannot be mapped); nccl = NCCL send/recv "
                         "of the packed faces on a second stream; peer = stores into the neighbour's memory inside the step kernel")
    ap.add_argument("--nccl-max-ctas", type=int, default=0,
                    help="d3q19_config.nccl_max_ctas: cap on the CTAs of NCCL's face send/recv (0 = NCCL's own choice, the default)")
    ap.add_argument("--halo-split-min", type=int, default=0,
                    help="d3q19_config.halo_split_min: peer-memory transports run slabs at least this thick as boundary + interior "
                         "launches, thinner ones as one launch (0 = the library's default, 64)")
    ap.add_argument("--cpu-steps", type=int, default=20,
                    help="timed steps of the CPU arm (20 steps of 512x256x256 = about 10 s on 16 cores)")
    ap.add_argument("--particles", type=int, default=0,
                    help="configs[4]: N finite-size spheres (interpolated bounce-back, refill, momentum-exchange force); "
                         "a step is then links + collide-stream + IBB + lubrication + move + refill")
    ap.add_argument("--rad", type=float, default=15.0, help="sphere radius (var_inc.f90:67)")
    ap.add_argument("--device-init", action="store_true",
                    help="initvel+initpop on the device (implied by c4: the field does not fit a host staging copy)")
    return ap.parse_args()


def workload_dims(name):
    named = {"c1": (64, 32, 32), "c2": (512, 256, 256), "c4": (1024, 1024, 944)}
    if name in named:
        return named[name]
    nx, ny, nz = (int(t) for t in name.lower().split("x"))
    return nx, ny, nz


def kernel_source_hash():
    """sha256 over the sources that define the step kernels (what an ncu capture of k_step is valid for)"""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "d3q19-single-phase_b200", "csrc")
    for n in ("kernels.cuh", "collide.cuh", "lattice.cuh"):
        h.update(n.encode()); h.update(open(os.path.join(d, n), "rb").read())
    return h.hexdigest()


def shared_config(nx, ny, nz, nz_local, particles, rad):
    """the `config` object of BOTH arms (the driver compares them): what is computed, not how"""
    return {"workload": workload_name(nx, ny, nz), "per_gpu": "%dx%dx%d z-slab" % (nx, ny, nz_local), "precision": "fp64",
            "particles": ("%d moving spheres of radius %g: links, interpolated bounce-back, momentum-exchange force, "
                          "lubrication, move, refill every step; avedensity every 100 steps (main.f90:163-167)" % (particles, rad)) if particles else "none"}


def workload_name(nx, ny, nz):
    return ("D3Q19 MRT channel %dx%dx%d (nx x ny x nz, x wall-normal), turbulent set Re_tau=180, "
            "uniform body force, half-way bounce-back walls" % (nx, ny, nz))


# ---- clocks during the timed region (B200_PROFILING.md "clocks line") -------------------------
class ClockSampler:
    """SM clock, power and throttle reasons of one GPU, sampled IN PROCESS through NVML by a thread (every 5 ms) while the
    main thread sits in the library call: works under torchrun, needs no nvidia-smi start-up time, and a 30 ms timed
    region still gets several samples.  `nvidia-smi` is only the fallback when the NVML binding is missing."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.rows, self.th, self.stop_flag, self.t_mark, self.err = index, [], None, False, 0.0, None
        self.nv = self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remaps ordinals: find the NVML device of CUDA device `index` by its PCI bus id
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                toks = [t.strip() for t in vis.split(",") if t.strip()]
                if index < len(toks) and toks[index].isdigit():
                    phys = int(toks[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:                       # no NVML binding / no driver: say so in the record
            self.err, self.nv = "%s: %s" % (type(exc).__name__, exc), None

    def pin_to_gpu_numa_node(self):
        """bind this process to the CPUs next to its GPU (NVML's ideal affinity) BEFORE any host buffer is allocated, so
        that pinned staging memory is first-touched on the GPU's NUMA node; returns a short description"""
        if not self.nv:
            return "no NVML"
        try:
            self.nv.nvmlDeviceSetCpuAffinity(self.h)
            cpus = sorted(os.sched_getaffinity(0))
            return "process bound to the %d CPUs NVML lists for GPU %d (%d..%d)" % (len(cpus), self.index, cpus[0], cpus[-1])
        except Exception as exc:
            return "not bound (%s)" % exc

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = None
                self.rows.append((time.perf_counter(), float(mhz), int(rs), pw))
            except Exception as exc:
                self.err = "%s: %s" % (type(exc).__name__, exc)
                return
            time.sleep(0.005)

    def start(self):
        if self.nv:
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()

    def mark(self):
        """start of the timed region: earlier samples (warm-up) are dropped when later ones exist"""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["NVML unavailable: %s" % self.err]}
        self.stop_flag = True
        self.th.join(timeout=2)
        inside = [r for r in self.rows if r[0] >= self.t_mark]
        rows = inside if inside else self.rows[-3:]
        mask = 0
        for r in rows:
            mask |= r[2]
        pw = [r[3] for r in rows if r[3] is not None]
        return {"sm_mhz": statistics.median([r[1] for r in rows]) if rows else None, "sm_max_mhz": self.max_mhz,
                "samples": len(rows), "power_w": statistics.median(pw) if pw else None,
                "reasons": sorted(k for k, bit in self.REASONS.items() if mask & bit), "how": "NVML in process, 5 ms"}


# ---- the CPU arm: the restated reference on the host cores -----------------------------------------
def rank_grid(ncores, ny, nz):
    """nprocY x nprocZ with as many ranks as cores (para.f90:219-228 wants nprocY | nproc)."""
    npz = 1
    while npz * 2 <= ncores and nz % (npz * 2) == 0 and (npz * 2) ** 2 <= ncores * 2:
        npz *= 2
    npy = max(1, ncores // npz)
    while ny % npy:
        npy -= 1
    return npy, npz


def cpu_reference_run(nx, ny, nz, steps, warmup):
    """collision_MRT + macrovar per step like main.f90:157-161 on all host cores, one thread per
    MPI rank.  kind "reference": the reference's own Fortran machine-translated to C
    (oracle/_ref/libref_fast.so, built -O3 like the reference's Makefile:28 where /root/reference
    was mounted); kind "port": the hand restatement (oracle/d3q19_oracle.c) if that is missing."""
    from oracle import oracle as orc
    from oracle import ref
    ncores = os.cpu_count() or 1
    npy, npz = rank_grid(ncores, ny, nz)
    if ref.available(fast=True):
        w = ref.RefWorld(nx, ny, nz, nprocY=npy, nprocZ=npz, laminar=False, fast=True)
        w.run("initvel"); w.run("forcing"); w.run("initpop"); w.run("macrovar")     # main.f90:58-65,136
        w.loop("collision_mrt", "macrovar", warmup)
        t0 = time.perf_counter()
        w.loop("collision_mrt", "macrovar", steps)
        dt = time.perf_counter() - t0
        w.close()
        kind = "reference"
        what = ("the reference's collision.f90/para.f90/initial.f90 machine-translated to C (oracle/f90toc.py), gcc -O3 "
                "-march=x86-64-v3, in-process mini-MPI (no Fortran compiler / MPI in this image)")
    else:
        orc.lib(fast=True).orc_set_num_threads(ncores)
        w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=False, fast=True, nprocY=npy, nprocZ=npz)
        w.macrovar()
        for _ in range(warmup):
            w.collision_MRT(); w.macrovar()
        t0 = time.perf_counter()
        for _ in range(steps):
            w.collision_MRT(); w.macrovar()
        dt = time.perf_counter() - t0
        w.close()
        kind = "port"
        what = "C restatement of collision.f90 (oracle/d3q19_oracle.c) built -O3 -march=x86-64-v3 (no Fortran/MPI in this image)"
    mlups = nx * ny * nz * steps / dt / 1e6
    return {"value": mlups, "unit": "MLUPS", "cores": min(ncores, npy * npz), "kind": kind,
            "sample": "%dx%dx%d, %d warm-up + %d timed steps of collision_MRT+macrovar, %dx%d MPI ranks as threads; %s"
                      % (nx, ny, nz, warmup, steps, npy, npz, what),
            "ms_per_step": dt / steps * 1e3}


# ---- parity of the path being timed, inside the run that times it ---------------------------------------
def parity_check(pkg, rank, world, local_rank, fresh_nccl_id, connect, gather_ok, bcast, particles=True):
    """Before the timed region: (a) the committed golden vector of the REFERENCE (tests/golden/ref_slabs_*.npz, made by
    tests/golden/make_golden.py from the machine-translated Fortran; 16 z planes, 9 steps) run in STRICT arithmetic on
    this job's ranks -- one z-slab per GPU, the halo transport the bench uses -- must come back BIT FOR BIT, in both
    storage schemes, after 8 steps and after 9 (both in-place phases, the send-back after odd steps), with `macrovar`
    and the all-reduced `avedensity` (collision.f90:337-370, :500-501); (b) a moving-particle case on the N slabs must
    agree with the same case on ONE domain (run on rank 0's GPU) -- the particle path has no reference code (SURVEY
    fact 2), what is checked is that the decomposition does not change the answer.  Raises on mismatch."""
    import ctypes as C
    import numpy as np
    capi = pkg.capi
    path = os.path.join(ROOT, "tests", "golden", "ref_slabs_21x4x16_r1x4_s9.npz")
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    nx, ny, nz = meta["nx"], meta["ny"], meta["nz"]
    ustar = meta["overrides"]["ustar"]
    kw = dict(ustar=ustar, force_in_y=2.0 * ustar * ustar / float(nx), ystar=0.0036 / ustar)     # para.f90:64-66
    res = {"case": os.path.basename(path)[:-4], "ranks": world, "steps": meta["steps"], "schemes": [], "bit_exact": True}
    if nz < world:
        res.update(bit_exact=None, skipped="more ranks than z planes in the golden case")
        return res
    for name, scheme in (("aa", capi.SCHEME_AA), ("ab", capi.SCHEME_AB)):
        sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=local_rank, scheme=scheme,
                              math_mode=capi.MATH_STRICT, nccl_id=fresh_nccl_id(), **kw)
        for k, v in meta["scalars"].items():
            if k != "mrttype" and getattr(sim.v, k) != v:
                raise SystemExit("bench parity: scalar %s differs from the reference's para (%r vs %r)" % (k, getattr(sim.v, k), v))
        connect(sim)
        z0, z1 = sim.globalz, sim.globalz + sim.lz
        sim.FORCING()
        sim.upload_f(np.ascontiguousarray(z["f0"][z0:z1]))
        out = np.empty((sim.lz, ny, nx, 19))
        ok = True
        sim.run_device(meta["snap"])
        sim.download_f(out)
        ok &= bool(np.array_equal(out, z["f_snap"][z0:z1]))
        sim.run_device(meta["steps"] - meta["snap"])
        sim.download_f(out)
        ok &= bool(np.array_equal(out, z["f"][z0:z1]))
        sim.device_macrovar()
        for k in ("rho", "ux", "uy", "uz"):
            ok &= bool(np.array_equal(getattr(sim, k), z[k][z0:z1]))
        m, n = C.c_double(0), C.c_int64(0)
        capi.check(sim.L.d3q19_avedensity(sim.h, C.byref(m), C.byref(n)))
        mean_ref, scale = float(np.mean(z["rho"])), float(np.mean(np.abs(z["rho"])))
        ok &= n.value == nx * ny * nz and abs(m.value - mean_ref) <= 1e-12 * scale     # the sum is order dependent
        sim.close()
        ok = gather_ok(ok)
        res["schemes"].append({"scheme": name, "bit_exact": ok})
        res["bit_exact"] = bool(res["bit_exact"] and ok)
    if particles and world > 1:
        # three spheres, two of them cut by slab faces; 3 resting + 6 moving steps (links, IBB, force all-reduce,
        # lubrication, move, refill with its source exchange); N slabs against one domain
        nx, ny, nz, rad = 24, 20, 8 * world, 3.6
        U = dict(ustar=0.0025, ystar=0.0036 / 0.0025, force_in_y=2.0 * 0.0025 * 0.0025 / nx)
        pos = [[11.7, 1.2, 8.0 * world - 0.9], [8.3, 12.0, 8.1], [15.5, 8.4, 4.2]]
        vel = [[0.010, 0.020, -0.010], [0.0, 0.015, 0.0], [-0.005, 0.0, 0.012]]
        omg = [[1e-3, 0.0, 2e-3], [0.0, -1e-3, 0.0], [5e-4, 5e-4, 0.0]]

        def run_case(r, nr, dev, nid):
            sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=r, nranks=nr, device=dev, scheme=capi.SCHEME_AB,
                                  nccl_id=nid, ipart=True, **U)
            sim.FORCING()
            if nr > 1:
                connect(sim, particles=True)
            sim.init_channel_device(A9=0.3, noise_amp=1e-3 * sim.v.ustar, seed=777)
            sim.particles_init(pos, rad, vel, omg)
            for _ in range(3):
                sim.particle_step(move=False)
            for _ in range(6):
                sim.particle_step(move=True)
            out = np.empty((sim.lz, ny, nx, 19))
            sim.download_f(out)
            g = sim.get_particles()
            mask = sim.get_mask()
            span = (sim.globalz, sim.globalz + sim.lz)
            sim.close()
            return out, g, mask, span

        ref = run_case(0, 1, local_rank, None) if rank == 0 else None
        f_ref, g_ref, mask_ref = bcast(None if ref is None else (ref[0], ref[1], ref[2]))
        out, g, mask, (z0, z1) = run_case(rank, world, local_rank, fresh_nccl_id())
        fluid = mask_ref[z0:z1] < 0
        ferr = float(np.max(np.abs(out[fluid] - f_ref[z0:z1][fluid])) / np.max(np.abs(f_ref)))
        perr = float(np.max(np.abs(g["ypglb"] - g_ref["ypglb"])))
        herr = float(np.max(np.abs(g["fHIp"] - g_ref["fHIp"])) / np.max(np.abs(g_ref["fHIp"])))
        okp = bool(np.array_equal(mask, mask_ref[z0:z1])) and ferr < 1e-10 and perr < 1e-11 and herr < 1e-9
        okp = gather_ok(okp)
        res["particles"] = {"case": "3 moving spheres, %dx%dx%d, 9 steps, %d slabs vs 1 domain" % (nx, ny, nz, world),
                            "mask_bit_exact": bool(np.array_equal(mask, mask_ref[z0:z1])), "f_rel_err": ferr,
                            "position_err": perr, "force_rel_err": herr, "ok": okp}
        res["bit_exact"] = bool(res["bit_exact"] and okp)
    return res


def main():
    args = parse_args()
    # rank 0 prints exactly ONE line on stdout: keep NCCL's "NCCL version ..." banner off it
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nx, ny, nz_unit = workload_dims(args.workload)
    n_gpus = max(args.gpus, world)
    nz = nz_unit * n_gpus if (args.scaling == "weak" and n_gpus > 1) else nz_unit

    if args.impl == "reference":
        if rank != 0:
            return 0
        # bounded sample: the single-GPU workload (the per-GPU block), a few steps
        cpu_steps = max(1, min(args.steps, args.cpu_steps))
        cpu_warm = max(0, min(args.warmup, 5))          # a CPU warm-up step costs ~0.6 s: the driver's W (>= 3) is honoured up to 5
        res = cpu_reference_run(nx, ny, nz_unit, cpu_steps, cpu_warm)
        nz_local = nz // n_gpus if n_gpus > 1 else nz
        line = {
            "impl": "reference", "metric": "MLUPS (fp64)", "value": res["value"], "unit": "MLUPS",
            "n_gpus": n_gpus, "steps": cpu_steps, "warmup": cpu_warm, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": shared_config(nx, ny, nz, nz_local, args.particles, args.rad),
            "implementation": {"note": "CPU arm: the reference's own hot path on the host cores; bounded sample = the per-GPU "
                                       "block %dx%dx%d, %d step(s)" % (nx, ny, nz_unit, cpu_steps)},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry

    pkg = entry.load_package()
    capi = pkg.capi
    if not torch.cuda.is_available() or capi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    if capi.is_hostsim() and os.environ.get("D3Q19_HOSTSIM_BENCH_STRUCTURE_TEST") != "1":
        raise SystemExit("bench.py: D3Q19_LIB points at the tests' host-sim build; it is never benchmarked")
    torch.cuda.set_device(local_rank)
    # clocks are sampled in process (NVML); with several ranks on one host every rank also binds itself to the CPUs next
    # to its GPU before it allocates pinned staging memory (the e2e leg uploads 5 GB per rank at the same time)
    sampler = ClockSampler(local_rank)
    numa_note = sampler.pin_to_gpu_numa_node() if world > 1 else "single process: not bound"
    nccl_id = None
    if world > 1:
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.tensor(list(capi.nccl_unique_id()), dtype=torch.uint8)
        dist.broadcast(idt, src=0)
        nccl_id = bytes(idt.tolist())

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    scheme = {"aa": capi.SCHEME_AA, "ab": capi.SCHEME_AB, "auto": capi.SCHEME_AUTO}[args.scheme]
    math_mode = capi.MATH_FAST if args.math == "fast" else capi.MATH_STRICT
    if args.particles > 0 and args.halo == "peer":
        args.halo = "put"               # with particles: copy engines or NCCL (the refill sources and the forces use NCCL anyway)
    nodes_global = nx * ny * nz
    device_init = args.device_init or args.workload == "c4"
    do_e2e = not args.no_e2e and not device_init and args.particles == 0

    def fresh_nccl_id():
        if world == 1:
            return None
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.tensor(list(capi.nccl_unique_id()), dtype=torch.uint8)
        dist.broadcast(idt, src=0)
        return bytes(idt.tolist())

    def all_ok(ok):
        """True when every rank says ok (the ranks must take the same branch afterwards)"""
        if world == 1:
            return bool(ok)
        t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)


    def build_sim(halo_req, nccl_id):
        """the channel on this rank's slab with its synthetic initial state; returns (sim, halo actually in use)"""
        sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=local_rank, scheme=scheme,
                              math_mode=math_mode, nccl_id=nccl_id, overlap=not args.no_overlap, allocate_host=False,
                              nccl_max_ctas=args.nccl_max_ctas, halo_split_min=args.halo_split_min, ipart=args.particles > 0)
        halo = "none"
        if world > 1:
            halo = "nccl"
            if halo_req in ("peer", "put"):
                def allgather_bytes(b):
                    t = torch.tensor(list(b), dtype=torch.uint8)
                    out = [torch.zeros_like(t) for _ in range(world)]
                    dist.all_gather(out, t)
                    return [bytes(o.tolist()) for o in out]
                # collective: either every rank maps its neighbours or none does (d3q19_ipc_connect agrees by all-reduce)
                ok_ = sim.connect_halo(allgather_bytes, mode="put" if halo_req == "put" else "fused")
                halo = halo_req if ok_ else "nccl"
        # synthetic initial state (turbulent set: log-law + perturbation + seeded noise)
        if device_init:
            sim.FORCING()
            sim.init_channel_device(A9=0.3, noise_amp=1e-3 * sim.v.ustar, seed=54321)
        else:
            sim.allocarray(pinned=True)
            sim.initvel(A9=0.3)
            sim.add_hash_noise(1e-3 * sim.v.ustar, seed=54321)
            sim.FORCING()
            sim.initpop()
            sim.upload_f()
        if args.particles > 0:
            # spheres on a regular lattice with more than mingap clearance, released from rest
            pitch = 2.0 * args.rad + 8.0
            slots = [(pitch * (i + 0.5) + 2.0, pitch * (j + 0.5), pitch * (k + 0.5))
                     for k in range(int(nz // pitch)) for j in range(int(ny // pitch)) for i in range(int((nx - 4) // pitch))]
            if len(slots) < args.particles:
                raise SystemExit("bench: %d spheres of radius %g do not fit %dx%dx%d" % (args.particles, args.rad, nx, ny, nz))
            stride = len(slots) / float(args.particles)
            pos = np.array([slots[int(i * stride)] for i in range(args.particles)], dtype=np.float64)
            sim.particles_init(pos, args.rad)
        return sim, halo

    # ---- parity of what is about to be timed: the reference's golden vector on this job's ranks and transport ----------
    parity = None
    if not args.no_parity:
        def allgather_bytes_(b):
            t = torch.tensor(list(b), dtype=torch.uint8)
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return [bytes(o.tolist()) for o in out]

        def connect_(sim_, particles=False):
            if world > 1 and args.halo in ("peer", "put"):
                mode = "put" if (args.halo == "put" or particles) else "fused"
                if not sim_.connect_halo(allgather_bytes_, mode=mode):
                    # agreed by all ranks inside d3q19_ipc_connect: no peer memory on this box -> NCCL send/recv from here on
                    if rank == 0:
                        sys.stderr.write("bench: peer memory cannot be mapped between these GPUs, the faces travel by NCCL\n")
                    args.halo = "nccl"

        def bcast_(obj):
            if world == 1:
                return obj
            box = [obj]
            dist.broadcast_object_list(box, src=0)
            return box[0]

        parity = parity_check(pkg, rank, world, local_rank, fresh_nccl_id, connect_, all_ok, bcast_)
        if parity["bit_exact"] is False:
            if rank == 0:
                sys.stderr.write("bench: PARITY CHECK FAILED before the timed region: %s\n" % json.dumps(parity))
                print(json.dumps({"impl": "ours", "parity_check": parity, "error": "parity check failed; nothing was timed"}))
            if world > 1:
                dist.destroy_process_group()
            return 3

    sim, halo = build_sim(args.halo, nccl_id)
    args.scheme = "ab" if sim.counters()["scheme"] == capi.SCHEME_AB else "aa"     # what AUTO resolved to

    pstep = [0]

    def advance(n):
        if args.particles > 0:
            for _ in range(n):
                sim.particle_step(move=True)
                pstep[0] += 1
                if pstep[0] % 100 == 0:            # main.f90:163-167: with particles, avedensity every 100 steps
                    capi.check(sim.L.d3q19_macrovar(sim.h))
                    capi.check(sim.L.d3q19_avedensity(sim.h, None, None))
        else:
            sim.run_device(n)

    # ---- device-timed value -------------------------------------------------------------------
    if rank == 0:
        sampler.start()
    ok = True
    try:
        advance(args.warmup)
        sim.sync()
    except RuntimeError as exc:          # the halo watchdog (d3q19_sync): a neighbour's flag never arrived
        ok = False
        sys.stderr.write("bench[rank %d]: %s\n" % (rank, exc))
    if not all_ok(ok):
        if halo not in ("peer", "put"):
            raise SystemExit("bench: the warm-up failed")
        # the peer-memory halo does not work on this box: rebuild everything on the NCCL transport
        if rank == 0:
            sys.stderr.write("bench: peer-memory halo failed, falling back to NCCL send/recv\n")
        sim.close()
        sim, halo = build_sim("nccl", fresh_nccl_id())
        advance(args.warmup)
        sim.sync()
    c0 = sim.counters()
    barrier(); sim.sync()
    sampler.mark()
    sim.timer_start()
    advance(args.steps)
    ms = sim.timer_stop()
    sim.sync(); barrier()
    clocks = sampler.stop() if rank == 0 else None
    c1 = sim.counters()
    ms = max_over_ranks(ms)
    mlups = nodes_global * args.steps / (ms * 1e-3) / 1e6
    launches = (c1["step_kernels"] - c0["step_kernels"]) + (c1["other_kernels"] - c0["other_kernels"])

    # sanity: the field is finite and still a channel flow
    pr = sim.probe(nx // 2, ny // 2, max(1, sim.lz // 2))
    if not np.all(np.isfinite(pr)):
        raise SystemExit("bench: non-finite field after the timed region")

    # roofline of the dominant kernel: per-launch algorithmic bytes / mean launch duration
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    nodes_local = nx * ny * sim.lz
    achieved = BYTES_PER_NODE * nodes_local / (ms * 1e-3 / args.steps) / 1e9
    # DRAM bytes per launch of the same kernel at the same per-GPU size from the committed ncu --set full capture
    # (profiles/traffic.json, made by tools/summarize_ncu.py).  The capture names the library build it was taken on
    # (sha256 of libd3q19b200.so's kernels source set): null when this run's library is another build or the size was
    # not captured -- a stale number is worse than none.
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            ent = tj.get("%s_%dx%dx%d" % (args.scheme, nx, ny, sim.lz))
            if ent and ent.get("kernel_source_sha256") == kernel_source_hash():
                traffic = ent.get("gb_per_launch")
                traffic_src = "profiles/traffic.json <- %s (same kernel sources)" % ent.get("source")
            elif ent:
                traffic_src = "capture %s was taken on other kernel sources: not reported" % ent.get("source")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "bytes_per_node": BYTES_PER_NODE, "peak_source": peak_src,
                "kernel": "k_step<%s>" % ("AA even/odd" if args.scheme == "aa" else "AB pull")}

    # ---- end to end through the reference-facing interface, host buffers ------------------------
    e2e = None
    if do_e2e:
        sim.v.nsteps = args.steps
        sim.set_schedule()
        sim.host_f_changed()                       # host f is the initial state again -> upload inside the timed region
        sim.initpop()
        barrier(); sim.sync()
        t0 = time.perf_counter()
        d2h = 0
        t_first = None
        for sim.istep in range(1, args.steps + 1):
            sim.collision_MRT()                    # first call uploads f (H2D, pinned)
            if t_first is None:
                sim.sync()
                t_first = time.perf_counter() - t0 # upload of f + one step
            sim.macrovar()                         # shim policy: downloads rho,u on output steps + the last
            pr = sim.probe(nx // 2, ny // 2, max(1, sim.lz // 2))
            d2h += 32
            if sim.istep % sim.v.nflowout == 0 or sim.istep % sim.v.ndiag == 0 or sim.istep == args.steps:
                d2h += 4 * 8 * nodes_local
        sim.sync(); barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": nodes_global * args.steps / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": 19 * 8 * nodes_local / args.steps,
               "d2h_bytes_per_step": d2h / args.steps,
               "what": "upload f from pinned host + K x (collision_MRT; macrovar; probe) through the shim entry points",
               # where the time goes: the one upload of f (19 x 8 B per node over PCIe) against K steps
               "upload_s": max_over_ranks(t_first), "total_s": dt,
               "upload_gbs_per_gpu": 19 * 8 * nodes_local / max(t_first, 1e-9) / 1e9,
               "numa": numa_note}

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        sim.close()
        cpu = cpu_reference_run(nx, ny, nz_unit, args.cpu_steps, 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": "MLUPS (fp64)", "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": shared_config(nx, ny, nz, sim.lz, args.particles, args.rad),
            "implementation": {
                       "scheme": args.scheme, "math": args.math,
                       "parallelism": ("z-slab x%d, faces %s" % (world, {
                           "peer": "stored into the neighbour GPU's memory over NVLink inside the step kernel",
                           "put": "copied into the neighbour GPU's memory over NVLink by the copy engines on a second stream",
                           "nccl": "by NCCL send/recv" + (" (at most %d CTAs)" % args.nccl_max_ctas if args.nccl_max_ctas else "")}[halo])) if world > 1 else "1 GPU",
                       "l2": "populations %.2f GB per GPU >> 126 MB L2 (no flush needed)" % (c1["population_bytes"] / 1e9)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "parity_check": parity,
            "clocks": clocks, "impl": "ours",
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
