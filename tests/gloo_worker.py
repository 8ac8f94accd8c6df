"""world_size-2 CPU worker (gloo) for tests/test_gloo_slabs.py.

Exercises, without a GPU, the host-side logic every N>1 run depends on:
  * the z-slab partition (`slab`, para.f90:240-261) and ring neighbours (para.f90:266-267);
  * the ghost exchange protocol of collisionExchnge (collision.f90:337-370) carried over a REAL
    inter-process transport: each process owns one slab, runs the oracle's local sweep on it and
    moves the 5-population face buffers with torch.distributed send/recv (tags as MPI_ISEND/IRECV,
    collision.f90:351-355); the result must be bit-identical to the single-domain oracle;
  * the bootstrap plumbing bench.py / mgpu_worker.py use: a 128-byte id broadcast from rank 0,
    MAX / SUM reductions of scalars (avedensity, timing).
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

ZS_M, ZS_P, ZR_M, ZR_P = 12, 13, 14, 15      # orc_rank_array indices of the z 5-slot buffers
YS_M, YS_P, YR_M, YR_P = 8, 9, 10, 11


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    pkg = entry.load_package()
    ok = True

    # ---- bootstrap plumbing: 128-byte id from rank 0 --------------------------------------------
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.arange(128, dtype=torch.uint8)
    dist.broadcast(idt, src=0)
    ok &= bytes(idt.tolist()) == bytes(range(128))

    for nx, ny, nz in [(12, 6, 8), (9, 4, 7), (16, 3, 2)]:
        lz, gz = pkg.slab(nz, world, rank)
        sizes = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([lz, gz]))
        ok &= sum(int(s[0]) for s in sizes) == nz and all(int(sizes[r][1]) == sum(int(sizes[q][0]) for q in range(r))
                                                          for r in range(world))
        up, dn = (rank + 1) % world, (rank + world - 1) % world         # mzp, mzm

        # the whole decomposed run lives in every process, but each process advances ONLY its own rank
        w, p = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True, nprocY=1, nprocZ=world, ustar=0.0025)
        single, _ = orc.make_initial_state(nx, ny, nz, laminar=False, noise=True, ustar=0.0025)
        d = w.rank_dims(rank)
        ok &= (d["lz"], d["globalz"], d["mzp"], d["mzm"]) == (lz, gz, up, dn)
        L = w.L
        ybuf = lambda which: w.rank_array(rank, which, (d["lz"] + 2, nx, 5))
        zbuf = lambda which: w.rank_array(rank, which, (ny, nx, 5))
        for step in range(6):
            L.orc_rank_collision_local(w.h, rank)                      # local sweep + y pack
            ybuf(YR_M)[...] = ybuf(YS_P)                               # nprocY = 1: mym = myp = myself
            ybuf(YR_P)[...] = ybuf(YS_M)
            L.orc_rank_unpack_y_pack_z(w.h, rank)
            # MPI_ISEND(tmpzmS -> mzm, tag 1), MPI_ISEND(tmpzpS -> mzp, tag 0); IRECV mirror (collision.f90:351-355)
            sm, sp = torch.from_numpy(zbuf(ZS_M).copy()), torch.from_numpy(zbuf(ZS_P).copy())
            rm, rp = torch.empty_like(sm), torch.empty_like(sp)
            reqs = [dist.isend(sm, dn, tag=1), dist.isend(sp, up, tag=0),
                    dist.irecv(rm, dn, tag=0), dist.irecv(rp, up, tag=1)]
            for r in reqs:
                r.wait()
            zbuf(ZR_M)[...] = rm.numpy()
            zbuf(ZR_P)[...] = rp.numpy()
            L.orc_rank_unpack_z(w.h, rank)
            single.collision_MRT()
            mine = w.rank_array(rank, 0, (lz, ny, nx, 19))
            if not np.array_equal(mine, single.get_f()[gz:gz + lz]):
                print("rank %d: slab mismatch at step %d for %s" % (rank, step, (nx, ny, nz)), flush=True)
                ok = False
                break
        # scalar reductions as the library does them (sum + count -> mean; max of errors)
        local = torch.tensor([float(w.rank_array(rank, 0, (lz, ny, nx, 19)).sum()), float(lz * ny * nx)], dtype=torch.float64)
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
        ok &= int(local[1].item()) == nx * ny * nz
        ok &= abs(local[0].item() - float(single.get_f().sum())) <= 1e-12 * float(np.abs(single.get_f()).sum())
        w.close(); single.close()

    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("GLOO_SLABS_OK" if t.item() == 1 else "GLOO_SLABS_FAILED", flush=True)
    dist.destroy_process_group()
    return 0 if t.item() == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
