"""d3q19-single-phase_b200: B200 (sm_100a) implementation of the UDel-CFD D3Q19 Channel-Flow
time-step hot path behind the reference's own subroutine interface.

    csrc/      hand-written CUDA kernels + the C-ABI (include/d3q19_b200.h)
    fortran/   the ISO_C_BINDING replacement of collision.f90 (the drop-in shim)
    capi.py    ctypes binding of the C-ABI
    channel.py host mirror of the driver-facing interface (collision_MRT, macrovar, ...)
    saveload.py checkpoint files in the reference's Fortran-unformatted format

The directory name carries a hyphen (it is the reference's name); import it through
`__graft_entry__.load_package()` which registers it as `d3q19_single_phase_b200`.
"""
from . import capi                                    # noqa: F401
from . import saveload                                # noqa: F401
from .channel import ChannelFlow, VarInc, slab       # noqa: F401
