"""ctypes wrapper around the CPU oracle (oracle/d3q19_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package.  See
oracle/d3q19_oracle.h for what the oracle is (a restatement of
/root/reference/Channel-Flow/collision.f90, para.f90, initial.f90) and how far its
parity is pinned (bit for bit against the machine-translated reference, oracle/ref.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NPOP = 19


class Para(C.Structure):
    """Mirror of `orc_para` (oracle/d3q19_oracle.h); field order must match."""
    _fields_ = (
        [(n, C.c_int) for n in ("nx", "ny", "nz", "nprocY", "nprocZ", "laminar", "MRTtype", "ivel")]
        + [(n, C.c_double) for n in (
            "visc", "Rstar", "ustar", "force_in_y", "ystar", "force_mag", "rho0",
            "tau", "s1", "s2", "s4", "s9", "s10", "s13", "s16",
            "omegepsl", "omegepslj", "omegxx",
            "coef1", "coef2", "coef3", "coef4", "coef5", "coef3i", "coef4i",
            "val1", "val2", "val3", "val4", "val5", "val6", "val7", "val8", "val9",
            "val1i", "val2i", "val3i", "val4i", "val5i", "val6i", "val7i", "val8i", "val9i",
            "ww0", "ww1", "ww2", "pi", "pi2", "rhopart")]
        + [("ipart", C.c_int)]
        + [(n, C.c_int * NPOP) for n in ("cix", "ciy", "ciz", "ipopp")]
        + [("ipswap", C.c_int * 9), ("ipstay", C.c_int * 10)]
    )


def build(fast=False, quiet=True):
    """Compile the oracle with the recipe in oracle/Makefile (gcc, seconds)."""
    target = os.path.join(HERE, "liboracle_fast.so" if fast else "liboracle.so")
    src = os.path.join(HERE, "d3q19_oracle.c")
    hdr = os.path.join(HERE, "d3q19_oracle.h")
    if (not os.path.exists(target)
            or os.path.getmtime(target) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.run(["make", "-C", HERE, target], check=True,
                       stdout=subprocess.DEVNULL if quiet else None)
    return target


_libs = {}


def lib(fast=False):
    if fast not in _libs:
        L = C.CDLL(build(fast))
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        vp = C.c_void_p
        L.orc_para_init.argtypes = [C.POINTER(Para)] + [C.c_int] * 6
        L.orc_para_set_mrt.argtypes = [C.POINTER(Para)]
        L.orc_world_create.argtypes = [C.POINTER(Para)]
        L.orc_world_create.restype = vp
        L.orc_world_destroy.argtypes = [vp]
        L.orc_initvel.argtypes = [vp, C.c_double]
        for name in ("orc_initpop", "orc_forcing", "orc_collision_MRT", "orc_macrovar", "orc_rhoupdat",
                     "orc_deliver_y", "orc_deliver_z"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = None
        L.orc_vortcalc.argtypes = [vp, dp, dp, dp]
        L.orc_vortcalc.restype = None
        L.orc_sijstat.argtypes = [vp, dp]
        L.orc_sijstat.restype = None
        L.orc_avedensity.argtypes = [vp, C.POINTER(C.c_int64)]
        L.orc_avedensity.restype = C.c_double
        L.orc_gather_f.argtypes = [vp, dp]
        L.orc_scatter_f.argtypes = [vp, dp]
        L.orc_gather_scalar.argtypes = [vp, C.c_int, dp]
        L.orc_scatter_scalar.argtypes = [vp, C.c_int, dp]
        L.orc_scatter_ibnodes.argtypes = [vp, ip, ip]
        L.orc_rank_dims.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
        L.orc_rank_array.argtypes = [vp, C.c_int, C.c_int]
        L.orc_rank_array.restype = dp
        L.orc_rank_ibnodes.argtypes = [vp, C.c_int]
        L.orc_rank_ibnodes.restype = ip
        L.orc_world_para.argtypes = [vp]
        L.orc_world_para.restype = C.POINTER(Para)
        L.orc_world_set_para.argtypes = [vp, C.POINTER(Para)]
        L.orc_world_nproc.argtypes = [vp]
        L.orc_world_set_particles.argtypes = [vp, C.c_int, dp, dp, dp]
        for name in ("orc_rank_collision_local", "orc_rank_unpack_y_pack_z", "orc_rank_unpack_z"):
            getattr(L, name).argtypes = [vp, C.c_int]
            getattr(L, name).restype = None
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _libs[fast] = L
    return _libs[fast]


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_para(nx, ny, nz, laminar=True, nprocY=1, nprocZ=1, fast=False, **overrides):
    """para.f90:59-214 for run-time sizes.  `overrides` may set any orc_para field
    (e.g. MRTtype=3, visc=...); the relaxation set is re-derived afterwards."""
    p = Para()
    lib(fast).orc_para_init(C.byref(p), nx, ny, nz, int(bool(laminar)), nprocY, nprocZ)
    redo = False
    for k, v in overrides.items():
        setattr(p, k, v)
        redo |= k in ("MRTtype", "visc")
    if redo:
        lib(fast).orc_para_set_mrt(C.byref(p))
        for k, v in overrides.items():          # explicit s*/omeg* overrides win
            if k not in ("MRTtype", "visc"):
                setattr(p, k, v)
    return p


class World:
    """All ranks of an nprocY x nprocZ run of the reference, in one process.

    Global arrays use the reference's Fortran layouts seen from numpy as C-order
    f[iz, iy, ix, ip] and a[iz, iy, ix] (0-based)."""

    def __init__(self, para, fast=False):
        self.L = lib(fast)
        self.para = para
        self.h = self.L.orc_world_create(C.byref(para))
        self.nx, self.ny, self.nz = para.nx, para.ny, para.nz
        self.nproc = para.nprocY * para.nprocZ

    def close(self):
        if self.h:
            self.L.orc_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- the reference's subroutines -------------------------------------------------
    def initvel(self, A9=0.0):
        self.L.orc_initvel(self.h, A9)

    def initpop(self):
        self.L.orc_initpop(self.h)

    def FORCING(self):
        self.L.orc_forcing(self.h)

    def collision_MRT(self):
        self.L.orc_collision_MRT(self.h)

    def macrovar(self):
        self.L.orc_macrovar(self.h)

    def rhoupdat(self):
        self.L.orc_rhoupdat(self.h)

    def vortcalc(self):
        """saveload.f90:3929-4054 on the current ux,uy,uz -> global (ox, oy, oz)[iz,iy,ix]"""
        shape = (self.para.nz, self.para.ny, self.para.nx)
        o = [np.zeros(shape) for _ in range(3)]
        self.L.orc_vortcalc(self.h, *[a.ctypes.data_as(C.POINTER(C.c_double)) for a in o])
        return o

    def sijstat(self):
        """saveload.f90:2031-2091: Sij*Sij of the fluid nodes from f and the current rho,u -> global [iz,iy,ix] (0 at solid nodes)"""
        o = np.zeros((self.para.nz, self.para.ny, self.para.nx))
        self.L.orc_sijstat(self.h, o.ctypes.data_as(C.POINTER(C.c_double)))
        return o

    def avedensity(self):
        n = C.c_int64(0)
        m = self.L.orc_avedensity(self.h, C.byref(n))
        return m, n.value

    def set_para(self, para):
        self.para = para
        self.L.orc_world_set_para(self.h, C.byref(para))

    # --- global views ---------------------------------------------------------------
    def get_f(self):
        f = np.empty((self.nz, self.ny, self.nx, NPOP), dtype=np.float64)
        self.L.orc_gather_f(self.h, _dptr(f))
        return f

    def set_f(self, f):
        f = np.ascontiguousarray(f, dtype=np.float64)
        assert f.shape == (self.nz, self.ny, self.nx, NPOP)
        self.L.orc_scatter_f(self.h, _dptr(f))

    _which = {"rho": 0, "ux": 1, "uy": 2, "uz": 3, "fx": 4, "fy": 5, "fz": 6}

    def get(self, name):
        a = np.empty((self.nz, self.ny, self.nx), dtype=np.float64)
        self.L.orc_gather_scalar(self.h, self._which[name], _dptr(a))
        return a

    def set(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == (self.nz, self.ny, self.nx)
        self.L.orc_scatter_scalar(self.h, self._which[name], _dptr(a))

    def set_solid(self, ib, isn=None):
        """ib: (nz,ny,nx) int32, -1 fluid / >0 solid; isn: owning particle id (1-based)."""
        ib = np.ascontiguousarray(ib, dtype=np.int32)
        ip = C.POINTER(C.c_int32)
        isp = None
        if isn is not None:
            isn = np.ascontiguousarray(isn, dtype=np.int32)
            isp = isn.ctypes.data_as(ip)
        self.L.orc_scatter_ibnodes(self.h, ib.ctypes.data_as(ip), isp)

    def set_particles(self, ypglb, wp, omgp):
        ypglb, wp, omgp = (np.ascontiguousarray(a, dtype=np.float64) for a in (ypglb, wp, omgp))
        self.L.orc_world_set_particles(self.h, ypglb.shape[0], _dptr(ypglb), _dptr(wp), _dptr(omgp))

    # --- per-rank access (for the gloo test) -----------------------------------------
    def rank_dims(self, rid):
        out = (C.c_int * 9)()
        self.L.orc_rank_dims(self.h, rid, out)
        keys = ("lx", "ly", "lz", "globaly", "globalz", "mym", "myp", "mzm", "mzp")
        return dict(zip(keys, out))

    def rank_array(self, rid, which, shape):
        ptr = self.L.orc_rank_array(self.h, rid, which)
        return np.ctypeslib.as_array(ptr, shape=shape)


def synthetic_velocity(nx, ny, nz, ustar, seed=54321, amp=1e-3):
    """Seeded uniform noise of amplitude amp*ustar on all three components (SURVEY.md
    section 8(d) synthetic inputs; iflowseed=54321, var_inc.f90:63).  Not in the reference."""
    rng = np.random.default_rng(seed)
    return [amp * ustar * (2.0 * rng.random((nz, ny, nx)) - 1.0) for _ in range(3)]


def make_initial_state(nx, ny, nz, laminar=False, A9=0.0, noise=True, seed=54321, fast=False, nprocY=1, nprocZ=1,
                       **overrides):
    """main.f90:58-65: initvel; FORCING; initpop (+ optional seeded noise on u before initpop
    so that every population and halo value is distinct).  Returns (World, para)."""
    para = make_para(nx, ny, nz, laminar=laminar, nprocY=nprocY, nprocZ=nprocZ, fast=fast, **overrides)
    w = World(para, fast=fast)
    w.initvel(A9)
    if noise:
        for name, d in zip(("ux", "uy", "uz"), synthetic_velocity(nx, ny, nz, para.ustar, seed)):
            w.set(name, w.get(name) + d)
    w.FORCING()
    w.initpop()
    return w, para


# ---- saveload.f90 monitors restated in numpy (pinned to the translated reference's captured file output in
#      tests/test_oracle_ref.py; sums are order dependent, so these carry a 1e-12 tolerance, integers are exact) ----
def diag_line(w, ustar, solid=None):
    """diag (saveload.f90:1535-1640): fluid-node count, mean and rms velocity over the fluid nodes in wall units,
    the largest speed with the global 1-based location of its FIRST occurrence in the reference's k-j-i loop order
    (:1560-1573), solid volume fraction, max/min density fluctuation -- the numbers of one diag.dat line (:1662)."""
    ux, uy, uz, rho = (w.get(k) for k in ("ux", "uy", "uz", "rho"))
    fluid = np.ones(ux.shape, bool) if solid is None else ~solid
    nf = int(fluid.sum())
    um, vm, wm = (a[fluid].sum() / nf for a in (ux, uy, uz))
    rms = [np.sqrt((a[fluid] ** 2).sum() / nf - m * m) / ustar for a, m in ((ux, um), (uy, vm), (uz, wm))]
    vel = np.sqrt(ux * ux + uy * uy + uz * uz)
    vel[~fluid] = 0.0
    k, j, i = np.unravel_index(int(np.argmax(vel)), vel.shape)        # first occurrence in (z, y, x) order
    return dict(vmax=float(vel.max()), imout=i + 1, jmout=j + 1, kmout=k + 1, umean=um / ustar, vmean=vm / ustar,
                wmean=wm / ustar, urms=rms[0], vrms=rms[1], wrms=rms[2], volf=1.0 - nf / vel.size,
                rhomax=float(rho[fluid].max()), rhomin=float(rho[fluid].min()), nfluid=nf)


def plane_sums(w, solid=None):
    """the 11 per-x-plane sums of statistc / statistc2 (saveload.f90:1241-1266, :1393-1421) in d3q19_profiles
    order (ux,uy,uz,ux2,uy2,uz2,uxuy,uxuz,uyuz,rho,rho2) and the fluid-node count per plane"""
    ux, uy, uz, rho = (w.get(k) for k in ("ux", "uy", "uz", "rho"))
    m = np.ones(ux.shape) if solid is None else (~solid).astype(np.float64)
    q = [ux, uy, uz, ux * ux, uy * uy, uz * uz, ux * uy, ux * uz, uy * uz, rho, rho * rho]
    return np.stack([(a * m).sum(axis=(0, 1)) for a in q]), m.sum(axis=(0, 1))
