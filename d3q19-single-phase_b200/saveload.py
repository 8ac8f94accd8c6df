"""Checkpoint files in the reference's format (saveload.f90:196-231 savecntdflow, :296-332
loadcntdflow), so that a GPU run can restart from / hand back to
the CPU reference.  Host-side I/O, out of the hot path: the only coupling is that `f` must be
current on the host before a save (ChannelFlow.sync_f_to_host) and re-uploaded after a load
(ChannelFlow.host_f_changed).

A file is Fortran sequential-unformatted, one per rank:
    <dir>/endrunflow2D16x8.<7-digit step>.<3-digit rank>
    record 1:  istep, istat, imovie            (3 x int32)
    record 2:  f(0:18, lx, ly, lz)             (float64, column major = our f[iz,iy,ix,ip] bytes)
Each record is framed by 4-byte length markers; a record longer than 2147483639 bytes is split
into subrecords (leading marker negative when continued, trailing marker negative when it
continues a previous one) -- the ifort / gfortran convention.
"""
import os
import struct

import numpy as np

MAX_SUBRECORD = 2147483639


def filename(dirname, istep, rank):
    """saveload.f90:214-218"""
    return os.path.join(dirname, "endrunflow2D16x8.%07d.%03d" % (istep, rank))


def write_record(fh, payload, max_sub=MAX_SUBRECORD):
    view = memoryview(payload).cast("B")
    n, pos, first = len(view), 0, True
    while True:
        chunk = min(max_sub, n - pos)
        last = pos + chunk >= n
        fh.write(struct.pack("<i", chunk if last else -chunk))
        fh.write(view[pos:pos + chunk])
        fh.write(struct.pack("<i", chunk if first else -chunk))
        pos += chunk
        first = False
        if last:
            return


def read_record(fh):
    parts = []
    while True:
        head = struct.unpack("<i", fh.read(4))[0]
        data = fh.read(abs(head))
        if len(data) != abs(head):
            raise IOError("truncated Fortran record")
        tail = struct.unpack("<i", fh.read(4))[0]
        if abs(tail) != abs(head):
            raise IOError("Fortran record markers disagree: %d / %d" % (head, tail))
        parts.append(data)
        if head >= 0:
            return b"".join(parts)


def savecntdflow(sim, dirname, istat=0, imovie=0, max_sub=MAX_SUBRECORD):
    """saveload.f90:196-231.  Makes the host f current first (it lives on the GPU)."""
    f = sim.sync_f_to_host()
    istep = sim.v.istep0 + sim.v.nsteps                                  # :208
    os.makedirs(dirname, exist_ok=True)
    path = filename(dirname, istep, sim.rank)
    with open(path, "wb") as fh:
        write_record(fh, np.array([istep, istat, imovie], dtype="<i4").tobytes())     # :226
        write_record(fh, np.ascontiguousarray(f, dtype="<f8"), max_sub)                # :227
    return path


def loadcntdflow(sim, dirname, istpload):
    """saveload.f90:296-332: sets istep0 and f, then tells the library the host f changed."""
    path = filename(dirname, istpload, sim.rank)
    with open(path, "rb") as fh:
        istep0, istat, imovie = np.frombuffer(read_record(fh), dtype="<i4")
        raw = read_record(fh)
    f = np.frombuffer(raw, dtype="<f8")
    if f.size != sim.f.size:
        raise ValueError("%s holds %d values, this rank needs %d (lx,ly,lz = %d,%d,%d)"
                         % (path, f.size, sim.f.size, sim.lx, sim.ly, sim.lz))
    sim.f[...] = f.reshape(sim.f.shape)
    sim.v.istep0 = int(istep0)
    sim.istep = int(istep0)
    sim.host_f_changed()
    return int(istep0), int(istat), int(imovie)


def reslab(dirname_in, dirname_out, istep, nx, ny, nz, nranks_in, nranks_out, slab):
    """Re-decompose a checkpoint in z (the job of loadcntdflow_frmmore / _frmless,
    saveload.f90:336-434, for an arbitrary change of the slab count).  `slab(nz, n, r)` is the
    partition rule (channel.slab)."""
    planes = []
    head = None
    for r in range(nranks_in):
        lz, _ = slab(nz, nranks_in, r)
        with open(filename(dirname_in, istep, r), "rb") as fh:
            head = read_record(fh)
            planes.append(np.frombuffer(read_record(fh), dtype="<f8").reshape(lz, ny, nx, 19))
    full = np.concatenate(planes, axis=0)
    os.makedirs(dirname_out, exist_ok=True)
    for r in range(nranks_out):
        lz, gz = slab(nz, nranks_out, r)
        with open(filename(dirname_out, istep, r), "wb") as fh:
            write_record(fh, head)
            write_record(fh, np.ascontiguousarray(full[gz:gz + lz]))
