"""Turns the ncu artefacts a gpurun call brought back into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/r01b_launches.csv profiles/r01b_launches_summary.md
    python tools/summarize_ncu.py full gpurun_out/prof_r01b_ab.ncu-rep gpurun_out/prof_r01b_aa.ncu-rep profiles/r01b_ncu_full_summary.csv
"""
import collections
import csv
import json
import os
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers"]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[k].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[v].replace(",", ""))
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as fh:
        fh.write("# ncu launch list summary (`%s`)\n\n" % os.path.basename(src))
        fh.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        fh.write("| kernel | launches | total ms | mean us | share |\n|---|---|---|---|---|\n")
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.3f | %.1f | %.1f %% |\n" % (name, n, ns / 1e6, ns / n / 1e3, 100 * ns / total))
    print(open(dst).read())


def full(reps, dst):
    out = []
    traffic = {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = [hdr.index(w) for w in WANT if w in hdr]
        if not out:
            out.append(["capture"] + [hdr[i] for i in idx])
            out.append([""] + [units[i] for i in idx])
        for r in rows[2:]:
            out.append([os.path.basename(rep)] + [r[i] for i in idx])
            d = dict(zip(hdr, r))
            traffic[d["Kernel Name"] + " grid " + d["Grid Size"]] = {
                "dram_bytes_per_launch": (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])) * 1e9
                if units[hdr.index("dram__bytes_read.sum")] == "Gbyte" else None,
                "ms": float(d["gpu__time_duration.sum"])}
    with open(dst, "w", newline="") as fh:
        csv.writer(fh).writerows(out)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2:-1], sys.argv[-1])
