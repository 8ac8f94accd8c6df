"""Per-step timeline of the multi-GPU step (development aid for the strong-scaling gap, profiles/r01d_halo_transports.md).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/step_timeline.py \
        [--halo nccl|peer|put] [--scaling strong|weak] [--steps 60] [--size 512x256x256]

Every rank records four CUDA timing events per step (d3q19_trace_enable): before / after the boundary-plane launch,
after the interior launch, after the z-face exchange.  Rank 0 prints, per rank, the medians over the recorded steps of
    boundary   time of the boundary-plane launch
    interior   time of the interior launch
    exch_late  how long after the interior kernel the exchange finished (<= 0: fully hidden)
    gap        idle time on the compute stream between the end of one step and the start of the next
    step       start-to-start step time
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import ctypes as C  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--halo", default="nccl", choices=["nccl", "peer", "put"])
ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
ap.add_argument("--scheme", default="ab", choices=["aa", "ab"])
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--size", default="512x256x256")
ap.add_argument("--nccl-max-ctas", type=int, default=0)
ap.add_argument("--halo-split-min", type=int, default=0)
a = ap.parse_args()

pkg = entry.load_package()
capi = pkg.capi
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
nccl_id = None
if world > 1:
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.tensor(list(capi.nccl_unique_id()), dtype=torch.uint8)
    dist.broadcast(idt, src=0)
    nccl_id = bytes(idt.tolist())
nx, ny, nz = (int(t) for t in a.size.split("x"))
if a.scaling == "weak":
    nz *= world
sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, rank=rank, nranks=world, device=local, nccl_id=nccl_id,
                      scheme=capi.SCHEME_AA if a.scheme == "aa" else capi.SCHEME_AB, allocate_host=False,
                      nccl_max_ctas=a.nccl_max_ctas, halo_split_min=a.halo_split_min)
if world > 1 and a.halo != "nccl":
    def allgather_bytes(b):
        t = torch.tensor(list(b), dtype=torch.uint8)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [bytes(o.tolist()) for o in out]
    sim.connect_halo(allgather_bytes, mode="put" if a.halo == "put" else "fused")
sim.FORCING()
sim.init_channel_device(A9=0.3, noise_amp=1e-3 * sim.v.ustar)
sim.run_device(20)
sim.sync()
if world > 1:
    dist.barrier()
capi.check(sim.L.d3q19_trace_enable(sim.h, a.steps))
sim.run_device(a.steps)
n = C.c_int32(0)
buf = (C.c_float * (4 * a.steps))()
capi.check(sim.L.d3q19_trace_fetch(sim.h, C.byref(n), buf, a.steps))
t = np.array(buf[:4 * n.value], dtype=np.float64).reshape(n.value, 4) * 1e3          # microseconds
res = dict(rank=rank,
           boundary=float(np.median(t[:, 1] - t[:, 0])), interior=float(np.median(t[:, 2] - t[:, 1])),
           exch_late=float(np.median(t[:, 3] - t[:, 2])), gap=float(np.median(t[1:, 0] - t[:-1, 2])),
           step=float(np.median(np.diff(t[:, 0]))))
rows = [res]
if world > 1:
    rows = [None] * world
    dist.all_gather_object(rows, res)
if rank == 0:
    print(json.dumps(dict(halo=a.halo, nccl_max_ctas=a.nccl_max_ctas, halo_split_min=a.halo_split_min, scaling=a.scaling, scheme=a.scheme,
                          size=[nx, ny, nz], gpus=world, unit="us", ranks=rows)))
sim.close()
if world > 1:
    dist.destroy_process_group()
