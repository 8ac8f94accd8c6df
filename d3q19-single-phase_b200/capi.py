"""ctypes binding of include/d3q19_b200.h (libd3q19b200.so).

The product path is the CUDA library; if it is missing or there is no GPU the calls fail
loudly -- there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# D3Q19_LIB: another BUILD OF THE SAME LIBRARY -- kernel-variant sweeps (tools/kernel_sweep.py) and, on the GPU-less build box
# only, the tests' host-sim build of csrc/d3q19_api.cu (tests/host/make_hostsim.py; the package never builds or looks for it)
LIB_PATH = os.environ.get("D3Q19_LIB") or os.path.join(HERE, "libd3q19b200.so")

ABI_VERSION = 1
SCHEME_AA, SCHEME_AB, SCHEME_AUTO = 0, 1, 2
MATH_FAST, MATH_STRICT = 0, 1
MACRO_MAIN, MACRO_PRERELAX, MACRO_EXTERNAL = 0, 1, 2
NPOP = 19
IPC_BYTES = 256
HALO_FUSED, HALO_PUT = 0, 1

# every symbol include/d3q19_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "d3q19_create", "d3q19_destroy", "d3q19_sync", "d3q19_last_error", "d3q19_nccl_unique_id", "d3q19_device_count",
    "d3q19_ipc_export", "d3q19_ipc_connect", "d3q19_set_halo_mode", "d3q19_upload_f", "d3q19_download_f", "d3q19_set_macro", "d3q19_download_macro",
    "d3q19_init_channel", "d3q19_set_force_uniform", "d3q19_set_force_field", "d3q19_forcingp", "d3q19_download_force_field",
    "d3q19_collide_stream", "d3q19_run", "d3q19_macrovar", "d3q19_rhoupdat", "d3q19_avedensity", "d3q19_probe",
    "d3q19_prerelax", "d3q19_set_solid_mask", "d3q19_set_particles", "d3q19_profiles", "d3q19_profiles2", "d3q19_diag",
    "d3q19_vortcalc", "d3q19_download_vort", "d3q19_sijstat", "d3q19_download_sij2",
    "d3q19_particles_init", "d3q19_beads_links", "d3q19_beads_collision", "d3q19_beads_lubforce", "d3q19_beads_move",
    "d3q19_beads_filling", "d3q19_particle_step", "d3q19_get_particles", "d3q19_get_links", "d3q19_get_mask",
    "d3q19_timer_start", "d3q19_timer_stop", "d3q19_get_counters", "d3q19_trace_enable", "d3q19_trace_fetch",
    "d3q19_shim_bind", "d3q19_shim_bind_arrays", "d3q19_shim_set_schedule", "d3q19_shim_forcing", "d3q19_shim_rhoupdat", "d3q19_shim_collision_mrt",
    "d3q19_shim_prerelax_state",
    "d3q19_shim_macrovar", "d3q19_shim_avedensity", "d3q19_shim_sync_f_to_host", "d3q19_shim_sync_f_to_device",
]


class Config(C.Structure):
    """Mirror of d3q19_config."""
    _fields_ = [
        ("abi_version", C.c_int32),
        ("lx", C.c_int32), ("ly", C.c_int32), ("lz", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("globalz", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("device", C.c_int32),
        ("scheme", C.c_int32), ("math", C.c_int32), ("ipart", C.c_int32), ("overlap", C.c_int32),
        ("nccl_max_ctas", C.c_int32), ("pf_blocks", C.c_int32), ("halo_timeout_s", C.c_int32),
        ("halo_split_min", C.c_int32), ("force_idx64", C.c_int32),
        ("s1", C.c_double), ("s2", C.c_double), ("s4", C.c_double), ("s9", C.c_double),
        ("s10", C.c_double), ("s13", C.c_double), ("s16", C.c_double),
        ("omegepsl", C.c_double), ("omegepslj", C.c_double), ("omegxx", C.c_double),
        ("rhopart", C.c_double),
        ("reserved_d", C.c_double * 5),
        ("nccl_id", C.c_ubyte * 128),
    ]


class ShimArrays(C.Structure):
    """Mirror of d3q19_shim_arrays."""
    _fields_ = [
        ("f", C.POINTER(C.c_double)),
        ("rho", C.POINTER(C.c_double)), ("ux", C.POINTER(C.c_double)),
        ("uy", C.POINTER(C.c_double)), ("uz", C.POINTER(C.c_double)),
        ("force_realx", C.POINTER(C.c_double)), ("force_realy", C.POINTER(C.c_double)),
        ("force_realz", C.POINTER(C.c_double)),
        ("ibnodes", C.POINTER(C.c_int32)), ("isnodes", C.POINTER(C.c_int32)),
        ("ndiag", C.c_int32), ("nflowout", C.c_int32), ("nsteps_total", C.c_int32), ("istep0", C.c_int32),
        ("ntime", C.c_int32), ("prerelax_maxiter", C.c_int32), ("rhoepsl", C.c_double),
    ]


class ParticleParams(C.Structure):
    """Mirror of d3q19_particle_params."""
    _fields_ = [("rad", C.c_double), ("rho0", C.c_double), ("mingap", C.c_double), ("mingap_w", C.c_double),
                ("stf0", C.c_double), ("stf1", C.c_double), ("stf0_w", C.c_double), ("stf1_w", C.c_double),
                ("fscale", C.c_double), ("gforce", C.c_double * 3), ("maxlink", C.c_int64)]


class D3Q19Error(RuntimeError):
    pass


_lib = None


def is_hostsim(L=None):
    """True when the loaded library is the tests' host-sim build (it alone exports d3q19_hostsim_marker)"""
    L = L if L is not None else _lib
    try:
        return L is not None and getattr(L, "d3q19_hostsim_marker") is not None
    except AttributeError:
        return False


def load():
    """dlopen the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise D3Q19Error(
            "libd3q19b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "this package has no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    if is_hostsim(L):
        # only the tests' host-sim build exports this marker (tests/host/make_hostsim.py): never silent
        import sys
        sys.stderr.write("d3q19_b200: %s is the HOST-SIM test build -- the lattice is stepped on the CPU; nothing it "
                         "prints is a GPU result or a timing\n" % LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
    L.d3q19_last_error.restype = C.c_char_p
    L.d3q19_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.d3q19_destroy.argtypes = [vp]
    L.d3q19_sync.argtypes = [vp]
    L.d3q19_nccl_unique_id.argtypes = [C.POINTER(C.c_ubyte)]
    L.d3q19_device_count.argtypes = [C.POINTER(C.c_int32)]
    L.d3q19_ipc_export.argtypes = [vp, C.POINTER(C.c_ubyte)]
    L.d3q19_ipc_connect.argtypes = [vp, C.POINTER(C.c_ubyte)]
    L.d3q19_set_halo_mode.argtypes = [vp, C.c_int32]
    L.d3q19_trace_enable.argtypes = [vp, C.c_int32]
    L.d3q19_trace_fetch.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_int32]
    L.d3q19_upload_f.argtypes = [vp, dp]
    L.d3q19_download_f.argtypes = [vp, dp]
    L.d3q19_set_macro.argtypes = [vp, dp, dp, dp, dp]
    L.d3q19_download_macro.argtypes = [vp, dp, dp, dp, dp]
    L.d3q19_init_channel.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int32]
    L.d3q19_set_force_uniform.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    L.d3q19_set_force_field.argtypes = [vp, dp, dp, dp]
    L.d3q19_collide_stream.argtypes = [vp, C.c_int32]
    L.d3q19_run.argtypes = [vp, C.c_int32]
    L.d3q19_macrovar.argtypes = [vp]
    L.d3q19_rhoupdat.argtypes = [vp]
    L.d3q19_avedensity.argtypes = [vp, dp, C.POINTER(C.c_int64)]
    L.d3q19_probe.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, dp]
    L.d3q19_prerelax.argtypes = [vp, C.c_double, C.c_int32, ip, dp]
    L.d3q19_set_solid_mask.argtypes = [vp, ip, ip]
    L.d3q19_set_particles.argtypes = [vp, C.c_int32, dp, dp, dp]
    L.d3q19_profiles.argtypes = [vp, dp]
    L.d3q19_profiles2.argtypes = [vp, dp]
    L.d3q19_diag.argtypes = [vp, C.c_double, dp]
    L.d3q19_forcingp.argtypes = [vp, C.c_int32, C.c_double]
    L.d3q19_download_force_field.argtypes = [vp, dp, dp, dp]
    L.d3q19_vortcalc.argtypes = [vp]
    L.d3q19_download_vort.argtypes = [vp, dp, dp, dp]
    L.d3q19_sijstat.argtypes = [vp]
    L.d3q19_download_sij2.argtypes = [vp, dp]
    i64p = C.POINTER(C.c_int64)
    L.d3q19_particles_init.argtypes = [vp, C.c_int32, C.POINTER(ParticleParams)]
    L.d3q19_beads_links.argtypes = [vp, i64p]
    L.d3q19_beads_collision.argtypes = [vp]
    L.d3q19_beads_lubforce.argtypes = [vp]
    L.d3q19_beads_move.argtypes = [vp]
    L.d3q19_beads_filling.argtypes = [vp, i64p]
    L.d3q19_particle_step.argtypes = [vp, C.c_int32]
    L.d3q19_get_particles.argtypes = [vp, dp, dp, dp, dp, dp]
    L.d3q19_get_links.argtypes = [vp, C.c_int64, ip, ip, ip, ip, ip, dp, i64p]
    L.d3q19_get_mask.argtypes = [vp, ip]
    L.d3q19_timer_start.argtypes = [vp]
    L.d3q19_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.d3q19_get_counters.argtypes = [vp, C.POINTER(C.c_int64)]
    L.d3q19_shim_bind.argtypes = [vp, C.POINTER(ShimArrays)]
    L.d3q19_shim_bind_arrays.argtypes = [vp, dp, dp, dp, dp, dp, dp, dp, dp, ip, ip] + [C.c_int32] * 7 + [C.c_double]
    L.d3q19_shim_set_schedule.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    L.d3q19_shim_forcing.argtypes = [vp, C.c_double, C.c_double]
    L.d3q19_shim_rhoupdat.argtypes = [vp]
    L.d3q19_shim_collision_mrt.argtypes = [vp]
    L.d3q19_shim_prerelax_state.argtypes = [vp, dp, ip]
    L.d3q19_shim_macrovar.argtypes = [vp, C.c_int32]
    L.d3q19_shim_avedensity.argtypes = [vp]
    L.d3q19_shim_sync_f_to_host.argtypes = [vp]
    L.d3q19_shim_sync_f_to_device.argtypes = [vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise D3Q19Error(load().d3q19_last_error().decode(errors="replace"))


def dptr(a):
    """double* of a C-contiguous float64 numpy array (or NULL for None)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need a C-contiguous float64 array"
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"], "need a C-contiguous int32 array"
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def device_count():
    n = C.c_int32(0)
    rc = load().d3q19_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def nccl_unique_id():
    buf = (C.c_ubyte * 128)()
    check(load().d3q19_nccl_unique_id(buf))
    return bytes(buf)
