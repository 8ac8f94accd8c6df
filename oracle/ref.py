"""ctypes wrapper around oracle/_ref/libref.so: the reference's own Fortran hot path,
machine-translated to C by oracle/f90toc.py at build time (see that file's header).

TEST INFRASTRUCTURE ONLY.  The library exists only where it was built from the mounted
reference (this container; it then travels to the GPU box as a prebuilt .so).  `available()`
says whether it is there; nothing here reads /root/reference at run time.

    w = RefWorld(nx, ny, nz, nprocY=2, nprocZ=2, laminar=False)   # para + allocarray on every rank
    w.run("initvel"); w.run("forcing"); w.run("initpop")           # main.f90:58-65
    w.run("collision_mrt"); w.run("macrovar")                      # main.f90:157-161
    f = w.get_f()                                                  # global f[iz,iy,ix,ip]
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KIND_R, KIND_I = 8, 4


class Arr(C.Structure):
    _fields_ = [("p", C.c_void_p), ("kind", C.c_int), ("rank", C.c_int), ("owned", C.c_int),
                ("lo", C.c_int * 4), ("n", C.c_int * 4)]


def lib_path(fast=False, dropin=False):
    return os.path.join(HERE, "_ref", "libref_b200.so" if dropin else ("libref_fast.so" if fast else "libref.so"))


def available(fast=False, dropin=False):
    return os.path.exists(lib_path(fast, dropin))


_libs = {}


def lib(fast=False, dropin=None):
    """dropin: path of the d3q19 library (libd3q19b200.so or the tests' host-sim build) -> the build in which the
    reference's collision.f90 is replaced by this repository's Fortran shim (oracle/shim2c.py); the d3q19_* entry
    points the shim binds are resolved from that library, loaded first and globally"""
    fast = ("dropin", os.path.abspath(dropin)) if dropin else fast
    if fast not in _libs:
        if dropin:
            C.CDLL(dropin, mode=C.RTLD_GLOBAL)
            L = C.CDLL(lib_path(dropin=True))
            L.ref_shim_handle.argtypes = [C.c_void_p]
            L.ref_shim_handle.restype = C.c_void_p
        else:
            L = C.CDLL(lib_path(fast))
        L.ref_world_create.argtypes = [C.c_int]
        L.ref_world_create.restype = C.c_void_p
        L.ref_world_destroy.argtypes = [C.c_void_p]
        L.ref_world_state.argtypes = [C.c_void_p, C.c_int]
        L.ref_world_state.restype = C.c_void_p
        L.ref_world_set_override.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.ref_world_run.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_world_loop.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        L.ref_scalar.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.ref_scalar.restype = C.c_void_p
        L.ref_capture_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.ref_capture_get.restype = C.c_int
        L.ref_capture_clear.argtypes = [C.c_void_p]
        L.ref_world_set_playback.argtypes = [C.c_void_p, C.c_int, C.c_long, C.POINTER(C.c_double)]
        L.ref_world_playback_left.argtypes = [C.c_void_p, C.c_int]
        L.ref_world_playback_left.restype = C.c_long
        L.ref_array.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_array.restype = C.POINTER(Arr)
        _libs[fast] = L
    return _libs[fast]


class RefWorld:
    """All MPI ranks of one run of the translated reference (one thread per rank)."""

    def __init__(self, nx, ny, nz, nprocY=1, nprocZ=1, laminar=True, fast=False, ipart=False, dropin=None, **overrides):
        self.L = lib(fast, dropin)
        self.nx, self.ny, self.nz = nx, ny, nz
        self.nproc = nprocY * nprocZ
        self.h = self.L.ref_world_create(self.nproc)
        # what the reference fixes at compile time / hard-codes (var_inc.f90:51, para.f90:59,219)
        ov = dict(nx7=nx + 1, nx=nx, ny=ny, nz=nz, laminarflow=int(bool(laminar)), nprocy=nprocY)
        ov.update({k.lower(): v for k, v in overrides.items()})
        for k, v in ov.items():
            self.override(k, v)
        for r in range(self.nproc):                     # MPI_COMM_RANK / MPI_COMM_SIZE, main.f90:27-28
            self.set_scalar("myid", r, rank=r)
            self.set_scalar("nproc", self.nproc, rank=r)
        self.run("module_init")
        self.run("para")
        if ipart:            # para.f90:332 hard-codes ipart = .false. (after the translated part of `para`)
            self.set_scalar("ipart", 1)
        self.run("allocarray")

    def close(self):
        if self.h:
            self.L.ref_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def override(self, name, value):
        if self.L.ref_world_set_override(self.h, name.lower().encode(), float(value)):
            raise RuntimeError("override table full / name too long: %s" % name)

    def run(self, name):
        if self.L.ref_world_run(self.h, name.lower().encode()):
            raise KeyError("no translated subroutine %r" % name)

    def shim_handle(self, rank=0):
        """drop-in build only: the d3q19 handle the Fortran shim created on `rank` (None before its first hot-path call)"""
        return self.L.ref_shim_handle(self.L.ref_world_state(self.h, rank))

    def loop(self, a, b, nsteps):
        self.L.ref_world_loop(self.h, a.lower().encode(), (b or "").lower().encode(), nsteps)

    # ---- values the translated code wrote to file units (write(unit, fmt) lists) -----------------
    def captured(self, unit=None, rank=0):
        """numbers handed to write(unit, ...) by `rank` since the last clear, in the order written"""
        n = self.L.ref_capture_get(self.h, rank, 0, None, None)
        units = (C.c_int * max(n, 1))()
        vals = (C.c_double * max(n, 1))()
        self.L.ref_capture_get(self.h, rank, n, units, vals)
        u = np.array(units[:n], dtype=np.int64)
        v = np.array(vals[:n], dtype=np.float64)
        return v if unit is None else v[u == unit]

    def set_playback(self, values, rank=0):
        """what the translated read(unit) statements of `rank` (loadcntdflow) will be handed, in order"""
        v = np.ascontiguousarray(values, dtype=np.float64)
        self.L.ref_world_set_playback(self.h, rank, v.size, v.ctypes.data_as(C.POINTER(C.c_double)))

    def playback_left(self, rank=0):
        return self.L.ref_world_playback_left(self.h, rank)

    def clear_captured(self):
        self.L.ref_capture_clear(self.h)

    # ---- reflection ---------------------------------------------------------------------------
    def _scalar_ptr(self, name, rank):
        kind = C.c_int(0)
        p = self.L.ref_scalar(self.L.ref_world_state(self.h, rank), name.lower().encode(), C.byref(kind))
        if not p:
            raise KeyError("no module scalar %r" % name)
        return C.cast(p, C.POINTER(C.c_double if kind.value == KIND_R else C.c_int))

    def scalar(self, name, rank=0):
        return self._scalar_ptr(name, rank)[0]

    def set_scalar(self, name, value, rank=None):
        for r in (range(self.nproc) if rank is None else [rank]):
            self._scalar_ptr(name, r)[0] = value

    def array(self, name, rank=0):
        """numpy view (no copy) of a module array of one rank; Fortran dims (d1,d2,..) appear
        as the C-order shape (.., d2, d1).  Also returns the Fortran lower bounds."""
        a = self.L.ref_array(self.L.ref_world_state(self.h, rank), name.lower().encode())
        if not a or not a.contents.p:
            raise KeyError("no allocated module array %r" % name)
        a = a.contents
        shape = tuple(a.n[d] for d in reversed(range(a.rank)))
        ct = C.c_double if a.kind == KIND_R else C.c_int
        buf = C.cast(a.p, C.POINTER(ct))
        return np.ctypeslib.as_array(buf, shape=shape), [a.lo[d] for d in range(a.rank)]

    # ---- global views over the rank grid ---------------------------------------------------------
    def _place(self, r):
        ly, lz = self.scalar("ly", r), self.scalar("lz", r)
        gy, gz = self.scalar("globaly", r), self.scalar("globalz", r)
        return slice(gz, gz + lz), slice(gy, gy + ly)

    def get_f(self):
        out = np.empty((self.nz, self.ny, self.nx, 19))
        for r in range(self.nproc):
            sz, sy = self._place(r)
            out[sz, sy] = self.array("f", r)[0]
        return out

    def set_f(self, f):
        for r in range(self.nproc):
            sz, sy = self._place(r)
            self.array("f", r)[0][...] = f[sz, sy]

    def get(self, name):
        out = np.empty((self.nz, self.ny, self.nx))
        for r in range(self.nproc):
            sz, sy = self._place(r)
            out[sz, sy] = self.array(name, r)[0]
        return out

    def sij2(self):
        """sijstat00's first loop nest (saveload.f90:2031-2091): Sij*Sij of every fluid node from the non-equilibrium
        moments of f and the rho,u arrays macrovar left; the routine keeps it in an automatic array, whose elements the
        translation hands to the capture buffer (unit 99).  Global [iz,iy,ix]; solid nodes hold 0."""
        self.clear_captured()
        self.run("sijstat00")
        out = np.zeros((self.nz, self.ny, self.nx))
        for r in range(self.nproc):
            sz, sy = self._place(r)
            lz, ly = sz.stop - sz.start, sy.stop - sy.start
            out[sz, sy] = self.captured(99, rank=r).reshape(lz, ly, self.nx)
        self.clear_captured()
        return out

    def set_solid(self, ib, isn=None):
        """ib, isn: global (nz,ny,nx) int arrays; fills the interior of the ghosted ibnodes of every rank."""
        for r in range(self.nproc):
            sz, sy = self._place(r)
            self.array("ibnodes", r)[0][1:-1, 1:-1, 1:-1] = ib[sz, sy]
            if isn is not None:
                self.array("isnodes", r)[0][...] = isn[sz, sy]

    def set(self, name, a):
        for r in range(self.nproc):
            sz, sy = self._place(r)
            self.array(name, r)[0][...] = a[sz, sy]
