"""Checkpoint format of the reference (saveload.f90:196-231, :296-332): record framing, names,
sub-record splitting, z re-decomposition.  CPU only (a stub stands in for the GPU object)."""
import struct

import numpy as np

import __graft_entry__ as entry

pkg = entry.load_package()
sl = pkg.saveload


class Stub:
    """What saveload needs from ChannelFlow."""

    class V:
        istep0, nsteps = 300, 700

    def __init__(self, lx, ly, lz, rank=0):
        self.v, self.rank, self.lx, self.ly, self.lz = self.V(), rank, lx, ly, lz
        self.f = np.random.default_rng(rank).normal(size=(lz, ly, lx, 19))
        self.changed = 0

    def sync_f_to_host(self):
        return self.f

    def host_f_changed(self):
        self.changed += 1


def test_file_name_and_record_framing(tmp_path):
    s = Stub(5, 4, 3, rank=7)
    path = sl.savecntdflow(s, str(tmp_path), istat=2, imovie=1)
    assert path.endswith("endrunflow2D16x8.0001000.007")        # istep0 + nsteps, saveload.f90:208,214-218
    raw = open(path, "rb").read()
    assert struct.unpack("<i", raw[:4])[0] == 12 and struct.unpack("<i", raw[16:20])[0] == 12
    assert struct.unpack("<3i", raw[4:16]) == (1000, 2, 1)
    nbytes = 19 * 5 * 4 * 3 * 8
    assert struct.unpack("<i", raw[20:24])[0] == nbytes and len(raw) == 20 + 8 + nbytes
    # payload is f(0:18,lx,ly,lz) column-major = f[iz,iy,ix,ip] row-major
    assert np.array_equal(np.frombuffer(raw[24:24 + nbytes], dtype="<f8").reshape(3, 4, 5, 19), s.f)


def test_roundtrip_and_subrecords(tmp_path):
    s = Stub(6, 5, 4)
    sl.savecntdflow(s, str(tmp_path), max_sub=1000)              # forces the > 2 GiB convention on a small file
    raw = open(sl.filename(str(tmp_path), 1000, 0), "rb").read()
    assert struct.unpack("<i", raw[20:24])[0] == -1000            # continued sub-record
    t = Stub(6, 5, 4)
    t.f[...] = 0
    assert sl.loadcntdflow(t, str(tmp_path), 1000) == (1000, 0, 0)
    assert np.array_equal(t.f, s.f) and t.v.istep0 == 1000 and t.changed == 1


def test_reslab(tmp_path):
    nx, ny, nz = 4, 3, 10
    full = np.random.default_rng(5).normal(size=(nz, ny, nx, 19))
    for r in range(4):
        lz, gz = pkg.slab(nz, 4, r)
        s = Stub(nx, ny, lz, rank=r)
        s.f = np.ascontiguousarray(full[gz:gz + lz])
        sl.savecntdflow(s, str(tmp_path / "in"))
    sl.reslab(str(tmp_path / "in"), str(tmp_path / "out"), 1000, nx, ny, nz, 4, 3, pkg.slab)
    for r in range(3):
        lz, gz = pkg.slab(nz, 3, r)
        t = Stub(nx, ny, lz, rank=r)
        sl.loadcntdflow(t, str(tmp_path / "out"), 1000)
        assert np.array_equal(t.f, full[gz:gz + lz])


class StubM(Stub):
    """with the macroscopic arrays saveprerelax writes"""

    def __init__(self, lx, ly, lz, rank=0):
        super().__init__(lx, ly, lz, rank)
        rng = np.random.default_rng(100 + rank)
        self.rho, self.ux, self.uy, self.uz = (rng.normal(size=(lz, ly, lx)) for _ in range(4))
        self.istep = 0


def test_initflow_and_prerelax_files(tmp_path):
    s = StubM(5, 4, 3, rank=12)
    path = sl.saveinitflow(s, str(tmp_path), istat=7)
    assert path.endswith("finit.012")                                # saveload.f90:113-114
    raw = open(path, "rb").read()
    assert struct.unpack("<3i", raw[:12]) == (4, 7, 4)                # record 1: istat
    nbytes = 19 * 5 * 4 * 3 * 8
    assert struct.unpack("<i", raw[12:16])[0] == nbytes and len(raw) == 12 + 8 + nbytes
    t = StubM(5, 4, 3, rank=12)
    t.f[...] = 0
    assert sl.loadinitflow(t, str(tmp_path)) == 7 and np.array_equal(t.f, s.f) and t.changed == 1
    # prerelax: (f, rho) share ONE record, (ux, uy, uz) the next (saveload.f90:67-69)
    path = sl.saveprerelax(s, str(tmp_path), istep=345)
    assert path.endswith("prerelax_01/finit.012")
    raw = open(path, "rb").read()
    assert struct.unpack("<3i", raw[:12]) == (4, 345, 4)
    n = 5 * 4 * 3
    assert struct.unpack("<i", raw[12:16])[0] == (19 * n + n) * 8
    u = StubM(5, 4, 3, rank=12)
    for a in (u.f, u.rho, u.ux, u.uy, u.uz):
        a[...] = 0
    assert sl.loadprerelax(u, str(tmp_path)) == 345
    for k in ("f", "rho", "ux", "uy", "uz"):
        assert np.array_equal(getattr(u, k), getattr(s, k)), k


def _write_frm(dirname, istp, rank, head, f):
    import os
    os.makedirs(dirname, exist_ok=True)
    with open(sl.frm_filename(dirname, istp, rank), "wb") as fh:
        sl.write_record(fh, np.array(head, dtype="<i4").tobytes())
        sl.write_record(fh, np.ascontiguousarray(f))


def test_reference_reslab_loaders(tmp_path):
    nx, ny, nz = 4, 3, 12
    full = np.random.default_rng(8).normal(size=(nz, ny, nx, 19))
    # files of a 6-rank run, read by a 2-rank run (iprocrate = 3): saveload.f90:336-385
    for r in range(6):
        _write_frm(str(tmp_path / "six"), 4200, r, (4200, 1, 2), full[2 * r:2 * r + 2])
    assert sl.frm_filename("d", 4200, 5).endswith("endrunflow.004200.005")
    for r in range(2):
        t = Stub(nx, ny, 6, rank=r)
        assert sl.loadcntdflow_frmmore(t, str(tmp_path / "six"), 4200, 3) == (4200, 1, 2)
        assert np.array_equal(t.f, full[6 * r:6 * r + 6]) and t.v.istep0 == 4200
    # files of a 2-rank run, read by a 6-rank run: saveload.f90:388-434
    for r in range(2):
        _write_frm(str(tmp_path / "two"), 4200, r, (4200, 0, 0), full[6 * r:6 * r + 6])
    for r in range(6):
        t = Stub(nx, ny, 2, rank=r)
        sl.loadcntdflow_frmless(t, str(tmp_path / "two"), 4200, 3)
        assert np.array_equal(t.f, full[2 * r:2 * r + 2])


def test_cpu_yz_decomposition_to_gpu_slabs(tmp_path, oracle):
    # a CPU run on nprocY x nprocZ = 3 x 2 ranks with uneven blocks (para.f90:229-262) -> 4 z-slabs
    nx, ny, nz, npy, npz = 5, 8, 7, 3, 2
    full = np.random.default_rng(9).normal(size=(nz, ny, nx, 19))
    w, p = oracle.make_initial_state(nx, ny, nz, laminar=True, noise=False, nprocY=npy, nprocZ=npz)
    for rid in range(npy * npz):
        d = w.rank_dims(rid)                                      # the oracle's para: extents and offsets per rank
        indy, indz = rid % npy, rid // npy
        assert (d["ly"], d["globaly"]) == sl.yz_block(ny, npy, indy)
        assert (d["lz"], d["globalz"]) == sl.yz_block(nz, npz, indz)
        s = Stub(nx, d["ly"], d["lz"], rank=rid)
        s.f = np.ascontiguousarray(full[d["globalz"]:d["globalz"] + d["lz"], d["globaly"]:d["globaly"] + d["ly"]])
        sl.savecntdflow(s, str(tmp_path / "cpu"))
    sl.from_yz_ranks(str(tmp_path / "cpu"), str(tmp_path / "gpu"), 1000, nx, ny, nz, npy, npz, 4, pkg.slab)
    for r in range(4):
        lz, gz = pkg.slab(nz, 4, r)
        t = Stub(nx, ny, lz, rank=r)
        assert sl.loadcntdflow(t, str(tmp_path / "gpu"), 1000)[0] == 1000
        assert np.array_equal(t.f, full[gz:gz + lz])
    w.close()


# ---- the reference's own writers (saveload.f90) executed: what they name the file and what they put into its records -----
def _reference_records(w, sub, unit=12, rank=0):
    """run a translated checkpoint writer and rebuild (file-name pieces, [record payload bytes]) from what it handed to its
    character assignment and its unformatted write statements (oracle/f90toc.py captures both)"""
    w.clear_captured()
    w.run(sub)
    name = w.captured(9900, rank=rank)
    start = int(np.max(np.nonzero(name == -2.0)[0]))             # the last character assignment is the file name
    pieces = name[start + 1:]
    text = "".join("<dir>" if c == -1.0 else chr(int(c)) for c in pieces)
    struct_, vals = w.captured(9500 + unit, rank=rank), w.captured(9000 + unit, rank=rank)
    recs, pos, i = [], 0, 0
    while i < len(struct_):
        assert struct_[i] == -1.0
        i += 1
        payload = b""
        while i < len(struct_) and struct_[i] != -1.0:
            kind, cnt = int(struct_[i]), int(struct_[i + 1])
            chunk = vals[pos:pos + cnt]
            payload += chunk.astype("<i4" if kind == 4 else "<f8").tobytes()
            pos += cnt
            i += 2
        recs.append(payload)
    assert pos == len(vals)
    w.clear_captured()
    return text, recs


def _file_records(path):
    out = []
    with open(path, "rb") as fh:
        while True:
            try:
                out.append(sl.read_record(fh))
            except struct.error:
                return out


class RankView:
    """one rank of the translated reference seen through the attributes saveload.py reads"""

    def __init__(self, w, r):
        self.rank = r
        self.f = np.ascontiguousarray(w.array("f", r)[0])
        self.rho, self.ux, self.uy, self.uz = (np.ascontiguousarray(w.array(k, r)[0]) for k in ("rho", "ux", "uy", "uz"))
        self.lz, self.ly, self.lx = self.f.shape[:3]
        self.v = type("V", (), dict(istep0=int(w.scalar("istep0", r)), nsteps=int(w.scalar("nsteps", r))))()

    def sync_f_to_host(self):
        return self.f


def test_files_hold_what_the_reference_writes(tmp_path):
    # saveload.f90:196-231 (savecntdflow), :102-124 (saveinitflow), :50-73 (saveprerelax) machine-translated and RUN on a
    # 1 x 2 rank grid: file name and record payloads of saveload.py must be the reference's, byte for byte.  (The 4-byte
    # markers framing each record are the Fortran runtime's convention, not in the reference's source.)
    import pytest
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libref.so not built (no /root/reference here)")
    nx, ny, nz = 7, 6, 5
    w = ref.RefWorld(nx, ny, nz, nprocY=1, nprocZ=2, laminar=False, a9=0.3, ustar=0.0025)
    w.run("initvel"); w.run("forcing"); w.run("initpop"); w.run("macrovar")
    w.loop("collision_mrt", "macrovar", 3)
    w.set_scalar("istep0", 1200); w.set_scalar("nsteps", 345); w.set_scalar("istat", 4); w.set_scalar("imovie", 2)
    for r in (0, 1):
        view = RankView(w, r)
        # continued-run file
        text, recs = _reference_records(w, "savecntdflow", rank=r)
        path = sl.savecntdflow(view, str(tmp_path), istat=4, imovie=2)
        assert "<dir>" + path[len(str(tmp_path)) + 1:] == text                      # endrunflow2D16x8.0001545.00r
        assert _file_records(path) == recs and len(recs) == 2
        assert np.frombuffer(recs[0], dtype="<i4").tolist() == [1545, 4, 2]
        # initial-flow file
        text, recs = _reference_records(w, "saveinitflow", unit=10, rank=r)
        path = sl.saveinitflow(view, str(tmp_path), istat=4)
        assert "<dir>" + path[len(str(tmp_path)) + 1:] == text                      # finit.00r
        assert _file_records(path) == recs and len(recs) == 2
        # pre-relaxation file: (f, rho) in ONE record, then (ux, uy, uz)
        w.set_scalar("istep", 77)
        text, recs = _reference_records(w, "saveprerelax", unit=10, rank=r)
        path = sl.saveprerelax(view, str(tmp_path), 77)
        assert "<dir>" + path[len(str(tmp_path)) + 1:] == text                      # prerelax_01/finit.00r
        assert _file_records(path) == recs and len(recs) == 3
    w.close()
