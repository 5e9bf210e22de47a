#!/usr/bin/env python
"""Summarise an ncu report: per kernel the headline metrics, stall-reason totals and the hottest
SASS instructions.  Usage: tools/ncu_hot.py report.ncu-rep [top-n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units = raw[0], raw[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
names = []
for r in raw[2:]:
    kn = r[hdr.index("Kernel Name")]
    names.append(kn)
    print("==", kn)
    for w in want:
        if w in hdr:
            print("   %-62s %16s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))

seen = set()
for kn in names:
    short = kn.split("(")[0].split("<")[0].split()[-1]
    if short in seen:
        continue
    seen.add(short)
    src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--kernel-name", "regex:" + short))))
    if len(src) < 3:
        continue
    h = src[1]
    rows = [r for r in src[2:] if len(r) == len(h)]
    si = h.index("# Samples")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = {h[i]: 0 for i in stall_cols}
    for r in rows:
        for i in stall_cols:
            try:
                tot[h[i]] += int(r[i])
            except ValueError:
                pass
    allsamp = sum(int(r[si]) for r in rows if r[si].isdigit()) or 1
    print("\n## %s  samples=%d  instr=%d" % (short, allsamp, len(rows)))
    print("   stalls: " + ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / allsamp) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7]))
    rows.sort(key=lambda r: -(int(r[si]) if r[si].isdigit() else 0))
    for r in rows[:top]:
        best = max(stall_cols, key=lambda i: int(r[i]) if r[i].isdigit() else 0)
        print("   %5.1f%%  %-14s %s" % (100.0 * int(r[si]) / allsamp, h[best][6:], r[h.index("Source")][:110]))
