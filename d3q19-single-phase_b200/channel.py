"""Host-side mirror of the reference driver's interface for the Channel-Flow hot path.

The reference is a flat Fortran program: `main.f90` calls argument-less subroutines that talk
through `module var_inc`.  No Fortran compiler exists in this image, so this module plays the
part of the (intact) driver above the C-ABI: `ChannelFlow` owns the host arrays of var_inc
(same names, same layouts), and its methods carry the reference's subroutine names and call
the very entry points the replacement `collision.f90` shim binds
(`fortran/collision_b200.f90`): d3q19_shim_collision_mrt, d3q19_shim_macrovar, ...

    para / allocarray      para.f90:21-413, :418-503
    initvel / initpop      initial.f90:75-147, :19-46      (host, run once -- out of the hot path)
    FORCING                collision.f90:515-527
    rhoupdat               collision.f90:469-480
    collision_MRT          collision.f90:24-273
    macrovar               collision.f90:378-463
    avedensity             collision.f90:487-513
    prerelax / run         main.f90:70-90, :142-208

All compute goes through libd3q19b200.so on the GPU; there is no CPU fallback here.
Host arrays are numpy C-order views of the Fortran layouts: f[iz,iy,ix,ip], a[iz,iy,ix].
"""
import ctypes as C
import math

import numpy as np

from . import capi

CIX = np.array([0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0])
CIY = np.array([0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1])
CIZ = np.array([0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1])
IPOPP = np.array([0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15])


class VarInc:
    """The scalars of `module var_inc` that the path reads, as `para` sets them."""

    def __init__(self, nx, ny, nz, laminar=True, **overrides):
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.npop = 19
        self.rho0 = 1.0
        self.rhopart = 1.0                      # var_inc.f90:65
        self.pi = 4.0 * math.atan(1.0)          # var_inc.f90:71
        self.pi2 = 2.0 * self.pi
        self.ndiag, self.nflowout = 250, 100    # var_inc.f90:58-59
        self.ntime = 10000                      # var_inc.f90:59
        self.nsteps = 1000                      # para.f90:43
        self.istep0 = 0
        self.rhoepsl = 1.0e-05                  # para.f90:285
        self.ipart = False                      # para.f90:332
        self.laminar = bool(laminar)
        if not laminar:                         # para.f90:61-70
            self.visc = 0.0036
            self.Rstar = 180.0
            self.ustar = 2.0 * self.Rstar * self.visc / float(nx)
            self.force_in_y = 2.0 * self.rho0 * self.ustar * self.ustar / float(nx)
            self.ystar = self.visc / self.ustar
            self.force_mag = 1.0
            self.ivel = True
            self.MRTtype = 1
        else:                                   # para.f90:75-88
            self.Rstar = 20.0
            self.ustar = 0.05
            self.visc = 2.0 * self.ustar * float(nx) / self.Rstar
            self.force_in_y = 8.0 * self.visc * self.ustar / float(nx) ** 2
            self.ystar = self.visc / self.ustar
            self.force_mag = 1.0
            self.ivel = False
            self.MRTtype = 2
        for k, v in overrides.items():
            if not hasattr(self, k):
                raise AttributeError("unknown var_inc scalar %r" % k)
            setattr(self, k, v)
        self.set_mrt()
        for k, v in overrides.items():          # explicit rates win over the MRTtype presets
            if k.startswith("s") or k.startswith("omeg"):
                setattr(self, k, v)
        self.ww0, self.ww1, self.ww2 = 1.0 / 3.0, 1.0 / 18.0, 1.0 / 36.0   # para.f90:172-174

    def set_mrt(self):
        """para.f90:106-143"""
        self.tau = 3.0 * self.visc + 0.5
        self.s9 = 1.0 / self.tau
        self.s13 = self.s9
        if self.MRTtype == 1:
            self.s1, self.s2, self.s4, self.s10, self.s16 = 1.5, 1.4, 1.2, 1.4, 1.98
            self.omegepsl, self.omegepslj, self.omegxx = 0.0, -475.0 / 63.0, 0.0
        elif self.MRTtype == 2:
            self.s1 = self.s2 = self.s4 = self.s10 = self.s16 = self.s9
            self.omegepsl, self.omegepslj, self.omegxx = 3.0, -11.0 / 2.0, -1.0 / 2.0
        else:
            self.s1 = 1.8
            self.s2 = self.s10 = self.s16 = self.s1
            self.s4 = self.s9
            self.omegepsl, self.omegepslj, self.omegxx = 3.0, -11.0 / 2.0, -1.0 / 2.0


def slab(nz, nranks, rank):
    """z-slab of `rank`: para.f90:240-244 (uneven split) and :259-261 (global offset),
    with nprocY = 1 and nprocZ = nranks."""
    base, extra = (nz - nz % nranks) // nranks, nz - nranks * (nz // nranks)
    lz = base + 1 if rank < extra else base
    globalz = sum((base + 1 if r < extra else base) for r in range(rank))
    return lz, globalz


class ChannelFlow:
    """One rank (= one GPU) of a Channel-Flow run behind the reference's subroutine names."""

    def __init__(self, nx, ny, nz, laminar=True, rank=0, nranks=1, device=None, scheme=capi.SCHEME_AA,
                 math_mode=capi.MATH_FAST, nccl_id=None, overlap=True, allocate_host=True, nccl_max_ctas=0, pf_blocks=0,
                 halo_timeout_s=0, halo_split_min=0, force_idx64=False, **overrides):
        """the last five keywords are d3q19_config's tuning knobs (0 = the measured default); **overrides are scalars
        of `module var_inc` as `para` would set them"""
        self.v = VarInc(nx, ny, nz, laminar, **overrides)
        self.rank, self.nranks = int(rank), int(nranks)
        self.lx, self.ly = self.v.nx, self.v.ny
        self.lz, self.globalz = slab(self.v.nz, self.nranks, self.rank)
        self.istep = 0
        self.L = capi.load()
        cfg = capi.Config()
        cfg.abi_version = capi.ABI_VERSION
        cfg.lx, cfg.ly, cfg.lz = self.lx, self.ly, self.lz
        cfg.nx, cfg.ny, cfg.nz = self.v.nx, self.v.ny, self.v.nz
        cfg.globalz = self.globalz
        cfg.rank, cfg.nranks = self.rank, self.nranks
        cfg.device = int(device if device is not None else 0)
        cfg.scheme, cfg.math = int(scheme), int(math_mode)
        cfg.ipart = int(bool(self.v.ipart))
        cfg.overlap = int(bool(overlap))
        cfg.nccl_max_ctas, cfg.pf_blocks, cfg.halo_timeout_s = int(nccl_max_ctas), int(pf_blocks), int(halo_timeout_s)
        cfg.halo_split_min, cfg.force_idx64 = int(halo_split_min), int(bool(force_idx64))
        for k in ("s1", "s2", "s4", "s9", "s10", "s13", "s16", "omegepsl", "omegepslj", "omegxx", "rhopart"):
            setattr(cfg, k, float(getattr(self.v, k)))
        if self.nranks > 1:
            if nccl_id is None or len(nccl_id) != 128:
                raise ValueError("nranks > 1 needs the 128-byte ncclUniqueId of rank 0")
            C.memmove(cfg.nccl_id, bytes(nccl_id), 128)
        self.cfg = cfg
        self.h = C.c_void_p()
        capi.check(self.L.d3q19_create(C.byref(cfg), C.byref(self.h)))
        self._bound = False
        if allocate_host:
            self.allocarray()

    # ---- para.f90:418-503 ------------------------------------------------------------------
    def allocarray(self, pinned=False):
        shp = (self.lz, self.ly, self.lx)
        if pinned:
            import torch
            self._pin = [torch.zeros(shp + (19,), dtype=torch.float64).pin_memory()] + \
                        [torch.zeros(shp, dtype=torch.float64).pin_memory() for _ in range(7)]
            arrs = [t.numpy() for t in self._pin]
        else:
            arrs = [np.zeros(shp + (19,))] + [np.zeros(shp) for _ in range(7)]
        (self.f, self.rho, self.ux, self.uy, self.uz,
         self.force_realx, self.force_realy, self.force_realz) = arrs
        self.ibnodes = np.full((self.lz + 2, self.ly + 2, self.lx + 2), -1, dtype=np.int32)   # para.f90:442,447
        self.isnodes = None
        self._bind()

    def _bind(self):
        a = capi.ShimArrays()
        a.f = capi.dptr(self.f)
        a.rho, a.ux, a.uy, a.uz = (capi.dptr(x) for x in (self.rho, self.ux, self.uy, self.uz))
        a.force_realx, a.force_realy, a.force_realz = (capi.dptr(x) for x in
                                                       (self.force_realx, self.force_realy, self.force_realz))
        a.ibnodes = capi.iptr(self.ibnodes)
        a.isnodes = capi.iptr(self.isnodes)
        a.ndiag, a.nflowout = self.v.ndiag, self.v.nflowout
        a.nsteps_total, a.istep0 = self.v.nsteps, self.v.istep0
        a.ntime, a.prerelax_maxiter, a.rhoepsl = self.v.ntime, 15000, self.v.rhoepsl
        self._shim_arrays = a
        capi.check(self.L.d3q19_shim_bind(self.h, C.byref(a)))
        self._bound = True

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.d3q19_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- initial.f90:75-147 (host, once) ---------------------------------------------------
    def initvel(self, A9=0.0):
        v = self.v
        self.ux[...] = 0.0
        self.uy[...] = 0.0
        self.uz[...] = 0.0
        if not v.ivel:
            return
        nx = v.nx
        nxh = (nx + 1) // 2
        i = np.arange(1, nxh + 1, dtype=np.float64)
        yplus = (i - 0.5) / v.ystar
        prof = np.where(yplus < 10.8, yplus * v.ustar, (np.log(yplus) / 0.41 + 5.0) * v.ustar)
        self.uy[:, :, :nxh] = prof
        self.uy[:, :, nx - 1 - np.arange(nxh)] = prof          # uy(nx+1-i) = uy(i)
        if A9 != 0.0:
            alpha, beta9, cc = 1.0, 1.0, 60.0
            ccc1 = -float(v.ny) / v.pi2 / alpha / v.ystar * A9 * v.ustar / cc / cc
            kk = np.arange(1, self.lz + 1) + self.globalz
            jj = np.arange(1, self.ly + 1)
            z9 = (v.pi2 * (kk - 0.5) / float(v.nz))[:, None, None]
            y9 = (v.pi2 * (jj - 0.5) / float(v.ny))[None, :, None]
            yp = yplus[None, None, :]
            ccc9 = np.exp(-yp / cc)
            ph = alpha * y9 + beta9 * z9
            du = ccc1 * yp * ccc9 * np.sin(ph)
            dv = (A9 * v.ustar * (1.0 - ccc9 - yp / cc * ccc9)) * np.cos(ph)
            mirror = nx - 1 - np.arange(nxh)
            self.uy[:, :, :nxh] += du
            np.add.at(self.uy, (slice(None), slice(None), mirror), du)
            self.ux[:, :, :nxh] += dv
            np.add.at(self.ux, (slice(None), slice(None), mirror), dv)

    # ---- initvel + initpop evaluated on the device (fields too large for the host) -----------
    def init_channel_device(self, A9=0.0, noise_amp=0.0, seed=54321):
        v = self.v
        capi.check(self.L.d3q19_init_channel(self.h, v.ustar, v.ystar, A9, noise_amp, seed, int(bool(v.ivel))))

    @staticmethod
    def splitmix_unit(seed, comp, node):
        """numpy mirror of the device noise generator (csrc/kernels.cuh splitmix_unit)."""
        with np.errstate(over="ignore"):
            z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (np.uint64(3) * node.astype(np.uint64) + np.uint64(comp + 1))
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) * 2.0 - 1.0

    def add_hash_noise(self, noise_amp, seed=54321):
        """Host twin of the noise init_channel_device adds (same global-node hash)."""
        v = self.v
        kz = (np.arange(self.lz) + self.globalz)[:, None, None]
        jy = np.arange(self.ly)[None, :, None]
        ix = np.arange(self.lx)[None, None, :]
        node = (ix + v.nx * (jy + v.ny * kz)).astype(np.uint64)
        for comp, a in enumerate((self.ux, self.uy, self.uz)):
            a += noise_amp * self.splitmix_unit(seed, comp, node)

    # ---- initial.f90:19-46 (host, once) ----------------------------------------------------
    def initpop(self):
        v = self.v
        usqr = self.ux * self.ux + self.uy * self.uy + self.uz * self.uz
        usqr = 1.5 * usqr
        self.rho[...] = 0.0
        rho = self.rho
        self.f[..., 0] = v.ww0 * (rho - usqr)
        for ip in range(1, 19):
            G = CIX[ip] * self.ux + CIY[ip] * self.uy + CIZ[ip] * self.uz
            ww = v.ww1 if ip <= 6 else v.ww2
            self.f[..., ip] = ww * (rho + 3.0 * G + 4.5 * G * G - usqr)
        capi.check(self.L.d3q19_shim_sync_f_to_device(self.h))

    def host_f_changed(self):
        """Tell the shim the host copy of f was rewritten (loadcntdflow, saveload.f90:296-332)."""
        capi.check(self.L.d3q19_shim_sync_f_to_device(self.h))

    def sync_f_to_host(self):
        """Make host f current (before savecntdflow / saveinitflow, saveload.f90:120,227)."""
        capi.check(self.L.d3q19_shim_sync_f_to_host(self.h))
        return self.f

    # ---- the subroutines main.f90 calls ----------------------------------------------------
    def FORCING(self):
        capi.check(self.L.d3q19_shim_forcing(self.h, self.v.force_in_y, self.v.force_mag))

    def FORCINGP(self, istep=None):
        """collision.f90:529-602 on the device for step `istep` (default: the driver's current istep)."""
        capi.check(self.L.d3q19_forcingp(self.h, int(self.istep if istep is None else istep), self.v.force_in_y))

    def download_force_field(self):
        o = [np.zeros((self.lz, self.ly, self.lx)) for _ in range(3)]
        capi.check(self.L.d3q19_download_force_field(self.h, *[capi.dptr(a) for a in o]))
        return o

    def rhoupdat(self):
        capi.check(self.L.d3q19_shim_rhoupdat(self.h))

    def collision_MRT(self):
        capi.check(self.L.d3q19_shim_collision_mrt(self.h))

    def macrovar(self):
        capi.check(self.L.d3q19_shim_macrovar(self.h, self.istep))

    def avedensity(self):
        capi.check(self.L.d3q19_shim_avedensity(self.h))

    def probe(self, ix, iy, iz):
        out = np.zeros(4)
        capi.check(self.L.d3q19_probe(self.h, ix, iy, iz, capi.dptr(out)))
        return out

    def diag(self):
        """saveload.f90:1507-1676 on the device; returns the fields of a diag.dat line."""
        out = np.zeros(14)
        capi.check(self.L.d3q19_diag(self.h, self.v.ustar, capi.dptr(out)))
        keys = ("vmax", "imout", "jmout", "kmout", "umean", "vmean", "wmean", "urms", "vrms", "wrms", "volf", "rhomax",
                "rhomin", "nfluid")
        d = dict(zip(keys, out))
        for k in ("imout", "jmout", "kmout", "nfluid"):
            d[k] = int(d[k])
        return d

    def profiles2(self):
        out = np.zeros((12, self.lx))
        capi.check(self.L.d3q19_profiles2(self.h, capi.dptr(out)))
        return out

    @staticmethod
    def statistc_rows(sums, count, ustar, ystar, with_volf=False, nynz=None):
        """What statistc / statistc2 do on rank 0 after their MPI_ALLREDUCEs (saveload.f90:1281-1338, :1446-1498):
        sums = the 11 plane sums in d3q19_profiles order, count = ny*nz (statistc) or the per-plane fluid count
        (statistc2).  Returns one row per x-plane in the column order of profiles.dat / profiles2.dat."""
        sums = np.asarray(sums, dtype=np.float64)
        lx = sums.shape[1]
        cnt = np.broadcast_to(np.asarray(count, dtype=np.float64), (lx,))
        u2, u4 = ustar * ustar, (ustar * ustar) * (ustar * ustar)
        vx, vy, vz = (sums[k] / cnt / ustar for k in (0, 1, 2))
        sqx, sqy, sqz = (sums[k] / cnt / u2 for k in (3, 4, 5))
        sxy, sxz, syz = (sums[k] / cnt / u2 for k in (6, 7, 8))
        sqx, sqy, sqz = sqx - vx ** 2, sqy - vy ** 2, sqz - vz ** 2
        sxz, sxy, syz = sxz - vx * vz, sxy - vx * vy, syz - vy * vz
        pr = sums[9] / cnt / u2
        prsq = sums[10] / cnt / u4 - pr * pr
        i = np.arange(1, lx + 1, dtype=np.float64)
        lxh = lx // 2                                            # var_inc.f90:53
        yplus = np.where(i > lxh, ((lx - i) + 0.5) / ystar, (i - 0.5) / ystar)
        cols = [i - 0.5, yplus, vx, vy, vz, sqx, sqy, sqz, sxz, sxy, syz, pr, prsq]
        if with_volf:
            volf = cnt / float(nynz)
            cols += [volf, 1.0 - volf]
        return np.stack(cols, axis=1)

    def statistc(self):
        """rows of profiles.dat (saveload.f90:1202-1342) from device-side plane sums"""
        return self.statistc_rows(self.profiles(), self.v.ny * self.v.nz, self.v.ustar, self.v.ystar)

    def statistc2(self):
        """rows of profiles2.dat (saveload.f90:1348-1502): fluid nodes only, with the solid volume fraction"""
        s = self.profiles2()
        return self.statistc_rows(s[:11], s[11], self.v.ustar, self.v.ystar, with_volf=True, nynz=self.v.ny * self.v.nz)

    def vortcalc(self):
        """saveload.f90:3929-4054 on the device: vorticity of the velocity field of the last macrovar;
        returns this rank's (ox, oy, oz)[iz, iy, ix] (the reference's ox,oy,oz(lx,ly,lz), var_inc.f90:140)."""
        o = [np.zeros((self.lz, self.ly, self.lx)) for _ in range(3)]
        capi.check(self.L.d3q19_vortcalc(self.h))
        capi.check(self.L.d3q19_download_vort(self.h, *[capi.dptr(a) for a in o]))
        return o

    def sijstat(self):
        """first loop nest of sijstat00 (saveload.f90:2031-2091) on the device: Sij*Sij of this rank's fluid nodes from the
        non-equilibrium moments of the current populations and the rho,u of the last macrovar; [iz, iy, ix]"""
        o = np.zeros((self.lz, self.ly, self.lx))
        capi.check(self.L.d3q19_sijstat(self.h))
        capi.check(self.L.d3q19_download_sij2(self.h, capi.dptr(o)))
        return o

    def profiles(self):
        out = np.zeros((11, self.lx))
        capi.check(self.L.d3q19_profiles(self.h, capi.dptr(out)))
        return out

    # ---- main.f90:70-90 through the intact-driver calls -------------------------------------
    def prerelax(self, allreduce_max=None, maxiter=15000, verbose=False):
        self.istep = 0
        while True:
            rhop = self.rho.copy()
            self.rhoupdat()
            self.collision_MRT()
            rhoerr = float(np.max(np.abs(self.rho - rhop)))
            rhoerrmax = allreduce_max(rhoerr) if allreduce_max else rhoerr
            if verbose:
                print(self.istep, rhoerrmax)
            if rhoerrmax <= self.v.rhoepsl or self.istep > maxiter:
                return self.istep, rhoerrmax
            self.istep += 1

    # ---- the same loop kept on the device -----------------------------
    def prerelax_device(self, maxiter=15000):
        capi.check(self.L.d3q19_upload_f(self.h, capi.dptr(self.f)))
        capi.check(self.L.d3q19_set_macro(self.h, capi.dptr(self.rho), capi.dptr(self.ux), capi.dptr(self.uy),
                                          capi.dptr(self.uz)))
        it, err = C.c_int32(0), C.c_double(0.0)
        capi.check(self.L.d3q19_prerelax(self.h, self.v.rhoepsl, maxiter, C.byref(it), C.byref(err)))
        capi.check(self.L.d3q19_download_macro(self.h, capi.dptr(self.rho), None, None, None))
        capi.check(self.L.d3q19_download_f(self.h, capi.dptr(self.f)))
        return it.value, err.value

    def set_schedule(self):
        """tell the shim the loop bounds / output cadence the download policy keys on (main.f90:142,171,184)"""
        v = self.v
        capi.check(self.L.d3q19_shim_set_schedule(self.h, v.ndiag, v.nflowout, v.nsteps, v.istep0))

    # ---- main.f90:142-208 -------------------------------------------------------------------
    def run(self, nsteps=None, on_step=None, time_bond=None, allreduce_max=None):
        """the time loop; `time_bond` (seconds) is the wall-clock budget of main.f90:197-206: every `ntime` steps the
        loop is left when it is spent (the shim makes rho,u current on those steps for `probe`).  The reference decides
        from the MPI_ALLREDUCE MAX of the elapsed time (main.f90:200-204) so that every rank leaves at the same step;
        with nranks > 1 `allreduce_max(float) -> float` must do the same, otherwise the ranks would part ways and the
        remaining ones hang in the halo exchange"""
        import time
        if time_bond is not None and self.nranks > 1 and allreduce_max is None:
            raise ValueError("run(time_bond=...) on %d ranks needs allreduce_max (main.f90:200-204)" % self.nranks)
        v = self.v
        nsteps = v.nsteps if nsteps is None else nsteps
        v.nsteps = nsteps
        self.set_schedule()
        t0 = time.perf_counter()
        for self.istep in range(v.istep0 + 1, v.istep0 + nsteps + 1):
            self.collision_MRT()
            self.macrovar()
            if v.ipart and self.istep % 100 == 0:
                self.avedensity()
            if on_step:
                on_step(self)
            if time_bond is not None and self.istep % v.ntime == 0:
                elapsed = time.perf_counter() - t0
                if allreduce_max is not None:
                    elapsed = allreduce_max(elapsed)
                if elapsed > time_bond:
                    break
        return self.istep

    # ---- particles (the reference's beads_* phases; device-side bookkeeping) -----------------------
    def particles_init(self, ypglb, rad, wp=None, omgp=None, rho0=1.0, mingap=3.0, mingap_w=3.0, stf0=0.025, stf1=0.002,
                       stf0_w=0.025, stf1_w=0.002, fscale=0.0, gforce=(0.0, 0.0, 0.0), maxlink=0):
        ypglb = np.ascontiguousarray(ypglb, dtype=np.float64).reshape(-1, 3)
        self.npart = ypglb.shape[0]
        prm = capi.ParticleParams(rad, rho0, mingap, mingap_w, stf0, stf1, stf0_w, stf1_w, fscale,
                                  (C.c_double * 3)(*gforce), maxlink)
        capi.check(self.L.d3q19_particles_init(self.h, self.npart, C.byref(prm)))
        self.set_particles(ypglb, wp, omgp)

    def set_solid_mask(self, ibnodes, isnodes=None):
        """Static solid mask as the reference keeps it: ibnodes (-1 fluid / >0 solid) and isnodes (owning
        particle, 1-based), both given here WITHOUT ghosts as [lz, ly, lx]; the C-ABI takes the reference's
        ghosted ibnodes(0:lx+1,0:ly+1,0:lz+1) (para.f90:442), so the ghost shell is added (fluid)."""
        if ibnodes is None:
            capi.check(self.L.d3q19_set_solid_mask(self.h, None, None))
            return
        gh = np.full((self.lz + 2, self.ly + 2, self.lx + 2), -1, dtype=np.int32)
        gh[1:-1, 1:-1, 1:-1] = ibnodes
        isn = None if isnodes is None else np.ascontiguousarray(isnodes, dtype=np.int32)
        capi.check(self.L.d3q19_set_solid_mask(self.h, capi.iptr(gh), None if isn is None else capi.iptr(isn)))

    def set_particles(self, ypglb, wp=None, omgp=None):
        npart = getattr(self, "npart", 0) or np.asarray(ypglb).reshape(-1, 3).shape[0]
        self.npart = npart
        z = np.zeros((self.npart, 3))
        a = [np.ascontiguousarray(z if t is None else t, dtype=np.float64).reshape(-1, 3) for t in (ypglb, wp, omgp)]
        capi.check(self.L.d3q19_set_particles(self.h, self.npart, capi.dptr(a[0]), capi.dptr(a[1]), capi.dptr(a[2])))

    def beads_links(self):
        n = C.c_int64(0)
        capi.check(self.L.d3q19_beads_links(self.h, C.byref(n)))
        return n.value

    def beads_collision(self):
        capi.check(self.L.d3q19_beads_collision(self.h))

    def beads_lubforce(self):
        capi.check(self.L.d3q19_beads_lubforce(self.h))

    def beads_move(self):
        capi.check(self.L.d3q19_beads_move(self.h))

    def beads_filling(self):
        n = C.c_int64(0)
        capi.check(self.L.d3q19_beads_filling(self.h, C.byref(n)))
        return n.value

    def particle_step(self, move=True):
        capi.check(self.L.d3q19_particle_step(self.h, int(bool(move))))

    def get_particles(self):
        out = {k: np.zeros((self.npart, 3)) for k in ("ypglb", "wp", "omgp", "fHIp", "torqp")}
        capi.check(self.L.d3q19_get_particles(self.h, *(capi.dptr(out[k]) for k in ("ypglb", "wp", "omgp", "fHIp", "torqp"))))
        return out

    def get_links(self, canonical=True):
        """the boundary links of this slab.  The device list is a SET (warps append their rows as they finish);
        canonical=True returns it sorted by (particle, z, y, x, direction) so that two lists can be compared"""
        n = C.c_int64(0)
        cap = 1 << 16
        while True:
            a = [np.zeros(cap, dtype=np.int32) for _ in range(5)]
            q = np.zeros(cap)
            rc = self.L.d3q19_get_links(self.h, cap, *(capi.iptr(t) for t in a), capi.dptr(q), C.byref(n))
            if rc == 0:
                break
            if n.value > cap:
                cap = int(n.value)
                continue
            capi.check(rc)
        m = n.value
        d = dict(x=a[0][:m], y=a[1][:m], z=a[2][:m], ip=a[3][:m], part=a[4][:m], q=q[:m])
        if canonical:
            o = np.lexsort((d["ip"], d["x"], d["y"], d["z"], d["part"]))
            d = {k: v[o] for k, v in d.items()}
        return d

    def get_mask(self):
        own = np.zeros((self.lz, self.ly, self.lx), dtype=np.int32)
        capi.check(self.L.d3q19_get_mask(self.h, capi.iptr(own)))
        return own

    # ---- halo in NVLink peer memory (collective; `allgather(bytes) -> [bytes per rank]`) ----------
    def connect_halo(self, allgather, mode=None):
        """mode: None (library default: fused), "fused" (stores inside the step kernel) or "put" (copy engines)"""
        if mode is not None:
            capi.check(self.L.d3q19_set_halo_mode(self.h, {"fused": capi.HALO_FUSED, "put": capi.HALO_PUT}[mode]))
        blob = (C.c_ubyte * capi.IPC_BYTES)()
        capi.check(self.L.d3q19_ipc_export(self.h, blob))
        blobs = allgather(bytes(blob))
        assert len(blobs) == self.nranks and all(len(b) == capi.IPC_BYTES for b in blobs)
        buf = (C.c_ubyte * (capi.IPC_BYTES * self.nranks)).from_buffer_copy(b"".join(blobs))
        rc = self.L.d3q19_ipc_connect(self.h, buf)
        if rc == 2:          # agreed by all ranks: peer memory is not available, the halo stays on NCCL
            return False
        capi.check(rc)
        return True

    # ---- raw C-ABI conveniences (tests, bench) ------------------------------------------------
    def upload_f(self, f=None):
        capi.check(self.L.d3q19_upload_f(self.h, capi.dptr(self.f if f is None else f)))

    def download_f(self, out=None):
        out = self.f if out is None else out
        capi.check(self.L.d3q19_download_f(self.h, capi.dptr(out)))
        return out

    def collide_stream(self, mode=capi.MACRO_MAIN):
        capi.check(self.L.d3q19_collide_stream(self.h, mode))

    def run_device(self, nsteps):
        capi.check(self.L.d3q19_run(self.h, nsteps))

    def device_macrovar(self, download=True):
        capi.check(self.L.d3q19_macrovar(self.h))
        if download:
            capi.check(self.L.d3q19_download_macro(self.h, capi.dptr(self.rho), capi.dptr(self.ux),
                                                   capi.dptr(self.uy), capi.dptr(self.uz)))

    def set_macro(self, rho=None, ux=None, uy=None, uz=None):
        capi.check(self.L.d3q19_set_macro(self.h, capi.dptr(rho), capi.dptr(ux), capi.dptr(uy), capi.dptr(uz)))

    def set_force_uniform(self, fx, fy, fz):
        capi.check(self.L.d3q19_set_force_uniform(self.h, fx, fy, fz))

    def set_force_field(self, fx, fy, fz):
        capi.check(self.L.d3q19_set_force_field(self.h, capi.dptr(fx), capi.dptr(fy), capi.dptr(fz)))

    def sync(self):
        capi.check(self.L.d3q19_sync(self.h))

    def timer_start(self):
        capi.check(self.L.d3q19_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0.0)
        capi.check(self.L.d3q19_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def counters(self):
        out = (C.c_int64 * 8)()
        capi.check(self.L.d3q19_get_counters(self.h, out))
        return dict(step_kernels=out[0], other_kernels=out[1], nccl_ops=out[2], steps=out[3],
                    population_bytes=out[4], phase=out[5], x_pitch=out[6], scheme=out[7])
