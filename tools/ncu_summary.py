#!/usr/bin/env python
"""Text summary of an .ncu-rep for profiles/: per launch the headline raw metrics, then the executed
instruction mix and warp-stall samples from the source page.
usage: python tools/ncu_summary.py REPORT.ncu-rep [units_per_launch] > profiles/NAME.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "smsp__inst_executed.sum"]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    print(f"# {rep}" + (f"   (units per launch: {units:.0f})" if units else ""))
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, unit = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"\n== launch {d.get('ID')}: {d.get('Kernel Name')}  block {d.get('Block Size')} grid {d.get('Grid Size')}")
        for k in KEYS:
            if k in d:
                print(f"   {k:66s} {d[k]:>16s} {unit[hdr.index(k)]}")
        if units and "gpu__time_duration.sum" in d:
            u = unit[hdr.index("gpu__time_duration.sum")]
            t = float(d["gpu__time_duration.sum"]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1e-9)
            print(f"   algorithmic 16 B/unit / duration (under ncu, cold)                  {16 * units / t / 1e9:16.1f} GB/s")
    rows = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    i, seen = 0, set()
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name, hdr = rows[i][1], rows[i + 1]
            ix = {h: k for k, h in enumerate(hdr)}
            stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            ops, stalls, total, samples = Counter(), Counter(), 0, 0
            j = i + 2
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                r = rows[j]
                if len(r) >= len(hdr):
                    src = r[ix["Source"]].split()
                    op = src[1] if src and src[0].startswith("@") and len(src) > 1 else (src[0] if src else "?")
                    parts = op.split(".")
                    op = parts[0] + ("." + parts[-1] if parts[0] in ("LDG", "STG", "LDS", "STS", "LDL", "STL") and len(parts) > 1 else "")
                    n = int(r[ix["Instructions Executed"]] or 0)
                    ops[op] += n
                    total += n
                    for h in stall_cols:
                        v = int(r[ix[h]] or 0)
                        stalls[h] += v
                        samples += v
                j += 1
            key = (name, total)
            if key not in seen and total:
                seen.add(key)
                print(f"\n== instruction mix: {name}")
                print(f"   warp instructions {total}" + (f" = {32 * total / units:.1f} thread-instructions per unit" if units else ""))
                for op, n in ops.most_common(16):
                    print(f"     {op:12s} {n:12d} {100 * n / total:5.1f}%" + (f" {32 * n / units:7.2f}/unit" if units else ""))
                print("   stall samples: " + ", ".join(f"{h[6:]} {100 * v / max(samples, 1):.0f}%" for h, v in stalls.most_common(8)))
            i = j
        else:
            i += 1


if __name__ == "__main__":
    main()
