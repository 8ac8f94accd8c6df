"""Runs a few main-loop steps of one storage scheme on 512x256x256 (or the given size) -- the
short command that `ncu --set full -k regex:k_step` wraps (B200_PROFILING.md).  Prints nothing
that is a benchmark number."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scheme", default="ab", choices=["aa", "ab"])
ap.add_argument("--size", default="512x256x256")
ap.add_argument("--steps", type=int, default=8)
a = ap.parse_args()
pkg = entry.load_package()
capi = pkg.capi
nx, ny, nz = (int(t) for t in a.size.split("x"))
sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, scheme=capi.SCHEME_AA if a.scheme == "aa" else capi.SCHEME_AB,
                      allocate_host=False)
sim.FORCING()
sim.init_channel_device(A9=0.3, noise_amp=1e-3 * sim.v.ustar)
sim.run_device(a.steps)
sim.sync()
print("ran %d steps, scheme %s" % (a.steps, a.scheme))
sim.close()
