"""profiles/traffic.json from ncu --set full captures of the step kernels: DRAM bytes per launch, with the hash of the kernel
sources the capture is valid for (bench.py reports `roofline.traffic` only when the hash still matches).

    python tools/make_traffic.py SIZE gpurun_out/prof_r02x_ab.ncu-rep gpurun_out/prof_r02x_aa.ncu-rep [--hash HASH]

--hash: the kernel-source hash of the snapshot the capture was taken on (default: the sources as they are now)."""
import csv
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)

args = sys.argv[1:]
khash = bench.kernel_source_hash()
if "--hash" in args:
    i = args.index("--hash")
    khash = args[i + 1]
    del args[i:i + 2]
size, reps = args[0], args[1:]
nx, ny, nz = (int(t) for t in size.split("x"))
out = {"_what": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel, ncu --set full --clock-control none, "
                "B200.  Algorithmic bytes per launch at %s: 304 B x %d nodes = %.4f GB." % (size, nx * ny * nz, 304e-9 * nx * ny * nz)}
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    rd, wr, tm, kn = (hdr.index(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "Kernel Name"))
    scale = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}
    gbs, mss, names = [], [], []
    for r in rows[2:]:
        gbs.append(float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]])
        mss.append(float(r[tm]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[tm]])
        names.append(r[kn].split("(")[0])
    scheme = "aa" if len(gbs) > 1 else "ab"
    out["%s_%s" % (scheme, size)] = {"gb_per_launch": round(sum(gbs) / len(gbs), 4), "per_launch_gb": [round(g, 4) for g in gbs],
                                     "ncu_ms": [round(m, 4) for m in mss], "kernel": names, "source": os.path.basename(rep),
                                     "kernel_source_sha256": khash}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
