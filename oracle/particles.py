"""ctypes wrapper of oracle/particles_oracle.c -- the CPU checker of the particle kernels.

TEST INFRASTRUCTURE ONLY, and PARITY UNPINNED: the reference snapshot does not contain its
particle library (see particles_oracle.c).  Arrays are global (undecomposed) and use the
reference layouts seen from numpy in C order: f[iz,iy,ix,ip], own[iz,iy,ix] (owning particle
id, 1-based, or -1), particle tables (npart,3)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Geom(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("rad", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        so, src = os.path.join(HERE, "libparticles.so"), os.path.join(HERE, "particles_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-C", HERE, so], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        dp, ip, gp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(Geom)
        L.po_build_mask.argtypes = [gp, C.c_int, dp, ip]
        L.po_build_links.argtypes = [gp, C.c_int, dp, ip, C.c_long, ip, ip, ip, ip, ip, dp]
        L.po_build_links.restype = C.c_long
        L.po_ibb.argtypes = [gp, dp, ip, C.c_long, ip, ip, ip, ip, ip, dp, C.c_int, dp, dp, dp, C.c_double, dp, dp]
        L.po_refill.argtypes = [gp, dp, ip, ip, dp, dp, dp]
        L.po_refill.restype = C.c_long
        L.po_lubforce.argtypes = [gp, C.c_int, dp] + [C.c_double] * 7 + [dp]
        L.po_move.argtypes = [gp, C.c_int, C.c_double, C.c_double, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def canon(links, select=None):
    """link list (optionally a boolean selection of it) sorted by (particle, z, y, x, direction): the CUDA path builds
    its list as a set, comparisons are made after this sort"""
    d = {k: (v if select is None else v[select]) for k, v in links.items()}
    o = np.lexsort((d["ip"], d["x"], d["y"], d["z"], d["part"]))
    return {k: v[o] for k, v in d.items()}


class Particles:
    """Particle state + the per-step operations, on one whole channel."""

    def __init__(self, nx, ny, nz, rad, ypglb, wp=None, omgp=None, rhopart=1.0, rho0=1.0,
                 mingap=3.0, mingap_w=3.0, stf0=0.025, stf1=0.002, stf0_w=0.025, stf1_w=0.002, fscale=0.0,
                 gforce=(0.0, 0.0, 0.0)):
        self.g = Geom(nx, ny, nz, rad)
        self.nx, self.ny, self.nz, self.rad = nx, ny, nz, rad
        self.ypglb = np.ascontiguousarray(ypglb, dtype=np.float64).reshape(-1, 3).copy()
        self.npart = self.ypglb.shape[0]
        z = lambda: np.zeros((self.npart, 3))
        self.wp = z() if wp is None else np.ascontiguousarray(wp, dtype=np.float64).copy()
        self.omgp = z() if omgp is None else np.ascontiguousarray(omgp, dtype=np.float64).copy()
        self.fHIp, self.torqp, self.flubp, self.forcepp, self.torqpp, self.thetap = z(), z(), z(), z(), z(), z()
        self.rho0 = rho0
        self.volp = 4.0 / 3.0 * (4.0 * np.arctan(1.0)) * rad ** 3           # para.f90:340
        self.amp = rhopart * self.volp                                       # :341
        self.aip = 0.4 * self.amp * rad ** 2                                 # :342
        self.lub = (mingap, mingap_w, stf0, stf1, stf0_w, stf1_w, fscale)
        self.gforce = np.array(gforce, dtype=np.float64)
        self.own = np.full((nz, ny, nx), -1, dtype=np.int32)
        self.own0 = self.own.copy()
        self.links = None

    def build_mask(self):
        self.own0 = self.own.copy()
        lib().po_build_mask(C.byref(self.g), self.npart, _d(self.ypglb), _i(self.own))
        return self.own

    def build_links(self, maxlink=None):
        maxlink = maxlink or int(8 * self.npart * 4 * np.pi * (self.rad + 1) ** 2) + 64
        a = [np.zeros(maxlink, dtype=np.int32) for _ in range(5)]
        q = np.zeros(maxlink)
        n = lib().po_build_links(C.byref(self.g), self.npart, _d(self.ypglb), _i(self.own), maxlink,
                                 _i(a[0]), _i(a[1]), _i(a[2]), _i(a[3]), _i(a[4]), _d(q))
        assert n <= maxlink
        self.links = dict(x=a[0][:n].copy(), y=a[1][:n].copy(), z=a[2][:n].copy(), ip=a[3][:n].copy(),
                          part=a[4][:n].copy(), q=q[:n].copy())
        return self.links

    def ibb(self, f):
        """f: canonical populations after collision_MRT with the solid nodes skipped (modified in place)."""
        k = self.links
        lib().po_ibb(C.byref(self.g), _d(f), _i(self.own), len(k["q"]), _i(k["x"]), _i(k["y"]), _i(k["z"]), _i(k["ip"]),
                     _i(k["part"]), _d(k["q"]), self.npart, _d(self.ypglb), _d(self.wp), _d(self.omgp), self.rho0,
                     _d(self.fHIp), _d(self.torqp))

    def lubforce(self):
        lib().po_lubforce(C.byref(self.g), self.npart, _d(self.ypglb), *self.lub, _d(self.flubp))

    def move(self):
        lib().po_move(C.byref(self.g), self.npart, self.amp, self.aip, _d(self.fHIp), _d(self.torqp), _d(self.flubp),
                      _d(self.forcepp), _d(self.torqpp), _d(self.gforce), _d(self.ypglb), _d(self.wp), _d(self.omgp),
                      _d(self.thetap))

    def refill(self, f):
        return lib().po_refill(C.byref(self.g), _d(f), _i(self.own0), _i(self.own), _d(self.ypglb), _d(self.wp),
                               _d(self.omgp))
