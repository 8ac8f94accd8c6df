"""Checkpoint files in the reference's format (saveload.f90:196-231 savecntdflow, :296-332
loadcntdflow, :102-193 saveinitflow / loadinitflow, :50-99 saveprerelax / loadprerelax, :336-434
loadcntdflow_frmmore / _frmless), so that a GPU run can restart from /
hand back to the CPU reference -- including a CPU run decomposed in y AND z (para.f90:219-262), whose
rank files `from_yz_ranks` re-assembles into z-slabs.  Host-side I/O, out of the hot path: the only
coupling is that `f` must be current on the host before a save (ChannelFlow.sync_f_to_host) and
re-uploaded after a load (ChannelFlow.host_f_changed).

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


# ---- init-flow and pre-relaxation files (saveload.f90:50-193) -------------------------------------------------------
def initflow_filename(dirname, rank, prerelax=False):
    """saveload.f90:113-114 `finit.<3-digit rank>`; :61-62 `prerelax_01/finit.<rank>`"""
    return os.path.join(dirname, "prerelax_01" if prerelax else "", "finit.%03d" % rank)


def saveinitflow(sim, dirname, istat=0):
    """saveload.f90:102-124 (main.f90:101, after the pre-relaxation): record 1 istat, record 2 f."""
    f = sim.sync_f_to_host()
    os.makedirs(dirname, exist_ok=True)
    path = initflow_filename(dirname, sim.rank)
    with open(path, "wb") as fh:
        write_record(fh, np.array([istat], dtype="<i4").tobytes())                     # :119
        write_record(fh, np.ascontiguousarray(f, dtype="<f8"))                         # :120
    return path


def loadinitflow(sim, dirname):
    """saveload.f90:127-193 (main.f90:110): f from `finit.<rank>`; returns istat."""
    with open(initflow_filename(dirname, sim.rank), "rb") as fh:
        istat = int(np.frombuffer(read_record(fh), dtype="<i4")[0])
        f = np.frombuffer(read_record(fh), dtype="<f8")
    if f.size != sim.f.size:
        raise ValueError("finit.%03d holds %d values, this rank needs %d" % (sim.rank, f.size, sim.f.size))
    sim.f[...] = f.reshape(sim.f.shape)
    sim.host_f_changed()
    return istat


def saveprerelax(sim, dirname, istep):
    """saveload.f90:50-73: record 1 istep, record 2 (f, rho) in ONE record, record 3 (ux, uy, uz)."""
    f = sim.sync_f_to_host()
    path = initflow_filename(dirname, sim.rank, prerelax=True)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as fh:
        write_record(fh, np.array([istep], dtype="<i4").tobytes())                     # :67
        write_record(fh, np.concatenate([np.ravel(f), np.ravel(sim.rho)]).astype("<f8"))            # :68
        write_record(fh, np.concatenate([np.ravel(a) for a in (sim.ux, sim.uy, sim.uz)]).astype("<f8"))   # :69
    return path


def loadprerelax(sim, dirname):
    """saveload.f90:76-99; returns istep.  The caller re-uploads rho,u if it continues the pre-relaxation."""
    with open(initflow_filename(dirname, sim.rank, prerelax=True), "rb") as fh:
        istep = int(np.frombuffer(read_record(fh), dtype="<i4")[0])
        a = np.frombuffer(read_record(fh), dtype="<f8")
        b = np.frombuffer(read_record(fh), dtype="<f8")
    nf, n = sim.f.size, sim.rho.size
    if a.size != nf + n or b.size != 3 * n:
        raise ValueError("prerelax file of rank %d does not match lx,ly,lz = %d,%d,%d" % (sim.rank, sim.lx, sim.ly, sim.lz))
    sim.f[...] = a[:nf].reshape(sim.f.shape)
    sim.rho[...] = a[nf:].reshape(sim.rho.shape)
    for k, arr in enumerate((sim.ux, sim.uy, sim.uz)):
        arr[...] = b[k * n:(k + 1) * n].reshape(arr.shape)
    sim.host_f_changed()
    return istep


# ---- the reference's own re-slab loaders (saveload.f90:336-434) -----------------------------------------------------
def frm_filename(dirname, istpload, rank):
    """saveload.f90:365-369 / :417-421: `endrunflow.<6-digit step>.<3-digit rank>` (no '2D16x8', six digits)"""
    return os.path.join(dirname, "endrunflow.%06d.%03d" % (istpload, rank))


def _read_cntd(path):
    with open(path, "rb") as fh:
        head = np.frombuffer(read_record(fh), dtype="<i4")
        f = np.frombuffer(read_record(fh), dtype="<f8")
    return head, f


def loadcntdflow_frmmore(sim, dirname, istpload, iprocrate):
    """saveload.f90:336-385: this run has 1/iprocrate of the ranks that wrote the files; rank r reads the files of
    ranks r*iprocrate .. r*iprocrate + iprocrate-1, each holding lz/iprocrate planes."""
    lz9 = sim.lz // iprocrate
    if lz9 * iprocrate != sim.lz:
        raise ValueError("lz = %d is not a multiple of iprocrate = %d" % (sim.lz, iprocrate))
    head = None
    for ii in range(iprocrate):
        head, f9 = _read_cntd(frm_filename(dirname, istpload, sim.rank * iprocrate + ii))
        sim.f[ii * lz9:(ii + 1) * lz9] = f9.reshape(lz9, sim.ly, sim.lx, 19)            # :378
    sim.v.istep0 = int(head[0]); sim.istep = int(head[0])
    sim.host_f_changed()
    return tuple(int(t) for t in head)


def loadcntdflow_frmless(sim, dirname, istpload, iprocrate):
    """saveload.f90:388-434: this run has iprocrate times the ranks that wrote the files; rank r reads the file of
    rank r / iprocrate and keeps the (r mod iprocrate)-th part of its planes."""
    head, f9 = _read_cntd(frm_filename(dirname, istpload, sim.rank // iprocrate))
    ii = sim.rank % iprocrate
    f9 = f9.reshape(sim.lz * iprocrate, sim.ly, sim.lx, 19)
    sim.f[...] = f9[ii * sim.lz:(ii + 1) * sim.lz]                                       # :429
    sim.v.istep0 = int(head[0]); sim.istep = int(head[0])
    sim.host_f_changed()
    return tuple(int(t) for t in head)


# ---- from the reference's 2-D (y,z) decomposition to z-slabs ---------------------------------------------------------
def yz_block(n, nproc, ind):
    """para.f90:232-244 (extent) and :254-260 (global offset) of block `ind` of `nproc` along one direction"""
    base, extra = (n - n % nproc) // nproc, n - nproc * (n // nproc)
    ext = base + 1 if ind < extra else base
    off = sum((base + 1 if i < extra else base) for i in range(ind))
    return ext, off


def from_yz_ranks(dirname_in, dirname_out, istep, nx, ny, nz, nprocY, nprocZ, nranks_out, slab):
    """Re-assemble the checkpoint of a CPU run on nprocY x nprocZ ranks (rank = indz*nprocY + indy, para.f90:229-230)
    into `nranks_out` z-slab files.  Works plane-group by plane-group, so the full field is never held at once."""
    os.makedirs(dirname_out, exist_ok=True)
    head = None
    rows = []                                   # per indz: (globalz, lz, [memmap-free arrays per indy])
    for indz in range(nprocZ):
        lz, gz = yz_block(nz, nprocZ, indz)
        full = np.empty((lz, ny, nx, 19))
        for indy in range(nprocY):
            ly, gy = yz_block(ny, nprocY, indy)
            with open(filename(dirname_in, istep, indz * nprocY + indy), "rb") as fh:
                head = read_record(fh)
                full[:, gy:gy + ly] = np.frombuffer(read_record(fh), dtype="<f8").reshape(lz, ly, nx, 19)
        rows.append((gz, lz, full))
    for r in range(nranks_out):
        lz, gz = slab(nz, nranks_out, r)
        out = np.empty((lz, ny, nx, 19))
        for g0, l0, full in rows:
            a, b = max(gz, g0), min(gz + lz, g0 + l0)
            if a < b:
                out[a - gz:b - gz] = full[a - g0:b - g0]
        with open(filename(dirname_out, istep, r), "wb") as fh:
            write_record(fh, head)
            write_record(fh, out)
