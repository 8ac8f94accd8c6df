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
