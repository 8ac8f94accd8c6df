"""Times the step kernels of one library build (D3Q19_LIB=...) per kind: AA even, AA odd, AB.
Development tool for kernel-variant sweeps on the GPU box; prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
capi = pkg.capi
nx, ny, nz = (int(t) for t in (sys.argv[1] if len(sys.argv) > 1 else "512x256x256").split("x"))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
pf = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # d3q19_config.pf_blocks (0 = default 128, < 0 off)
rng = np.random.default_rng(0)
f0 = (1e-3 * rng.random((nz, ny, nx, 19))).astype(np.float64)
out = {"lib": os.path.basename(capi.LIB_PATH), "size": [nx, ny, nz], "pf_blocks": pf}
nodes = nx * ny * nz
for name, scheme in (("aa", capi.SCHEME_AA), ("ab", capi.SCHEME_AB)):
    sim = pkg.ChannelFlow(nx, ny, nz, laminar=False, scheme=scheme, allocate_host=False, pf_blocks=pf)
    sim.set_force_uniform(0.0, 1e-6, 0.0)
    sim.upload_f(f0)
    sim.run_device(6)
    sim.sync()
    t = {0: [], 1: []}
    for i in range(2 * reps):
        ph = sim.counters()["phase"]
        sim.timer_start()
        sim.run_device(1)
        t[ph].append(sim.timer_stop())
    if name == "aa":
        for ph, label in ((0, "aa_even"), (1, "aa_odd")):
            ms = float(np.median(t[ph]))
            out[label] = {"ms": round(ms, 4), "GBps": round(304.0 * nodes / ms / 1e6, 1)}
        ms = float(np.median(t[0]) + np.median(t[1])) / 2
        out["aa"] = {"ms": round(ms, 4), "MLUPS": round(nodes / ms / 1e3, 1), "GBps": round(304.0 * nodes / ms / 1e6, 1)}
    else:
        ms = float(np.median(t[0] + t[1]))
        out["ab"] = {"ms": round(ms, 4), "MLUPS": round(nodes / ms / 1e3, 1), "GBps": round(304.0 * nodes / ms / 1e6, 1)}
    sim.close()
print(json.dumps(out))
