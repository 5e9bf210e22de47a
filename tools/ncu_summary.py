#!/usr/bin/env python
"""Turn an ncu report + launch list into the tracked summaries under profiles/.

    tools/ncu_summary.py <tag>      reads gpurun_out/prof_<tag>.ncu-rep, launches_<tag>.csv, bench_<tag>.json
                                    writes profiles/<tag>_kernels.json, profiles/<tag>_summary.md,
                                           profiles/<tag>_launches.csv
"""
import csv
import io
import json
import os
import shutil
import statistics
import subprocess
import sys

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % tag)
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
M = {
    "time_us": "gpu__time_duration.sum", "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct", "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "registers": "launch__registers_per_thread", "grid": "launch__grid_size", "block": "launch__block_size",
    "warp_instructions": "smsp__inst_executed.sum",
}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3}
kernels = []
for r in rows[2:]:
    k = {"kernel": r[hdr.index("Kernel Name")]}
    for key, name in M.items():
        if name in hdr:
            i = hdr.index(name)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            k[key] = v * SCALE.get(units[i], 1.0)
    if "dram_read_bytes" in k:
        k["dram_traffic_bytes"] = k["dram_read_bytes"] + k.get("dram_write_bytes", 0.0)
    if "dram_pct" not in k:      # this ncu reports the DRAM share per direction
        tot = 0.0
        for name in ("dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed"):
            if name in hdr:
                try:
                    tot += float(r[hdr.index(name)].replace(",", ""))
                except ValueError:
                    pass
        k["dram_pct"] = tot
    if "lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed" in hdr:
        try:
            k["l2_tag_pct"] = float(r[hdr.index("lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed")].replace(",", ""))
        except ValueError:
            pass
    kernels.append(k)

launch_src = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
shares = {}
if os.path.exists(launch_src):
    shutil.copy(launch_src, os.path.join(out_dir, "%s_launches.csv" % tag))
    per = {}
    for r in csv.DictReader(l for l in open(launch_src) if l.startswith('"')):
        name = r["Kernel Name"].replace("void ", "")
        if not (name.startswith("kf_") or name.startswith("k_")):
            continue                      # torch's own setup kernels are not part of the step
        per.setdefault(r["Kernel Name"], []).append(float(r["Metric Value"]) / 1e3)
    tot = sum(statistics.mean(v) for v in per.values())
    shares = {k: {"mean_us": statistics.mean(v), "launches": len(v), "share": statistics.mean(v) / tot} for k, v in per.items()}

bench = None
bpath = os.path.join(ROOT, "gpurun_out", "bench_%s.json" % tag)
if os.path.exists(bpath):
    try:
        bench = json.loads(open(bpath).read().strip().splitlines()[-1])
    except Exception:
        bench = None

json.dump({"tag": tag, "kernels": kernels, "launch_list": shares, "bench": bench},
          open(os.path.join(out_dir, "%s_kernels.json" % tag), "w"), indent=1)

with open(os.path.join(out_dir, "%s_summary.md" % tag), "w") as f:
    f.write("# ncu summary `%s`\n\n" % tag)
    f.write("Source: `ncu --set full --clock-control none --import-source on` on one steady-state step "
            "(eager launches, single stream) and the launch list of the same command "
            "(`--metrics gpu__time_duration.sum`). ncu times are cold-cache and serialised: compare shares.\n\n")
    f.write("| kernel | time us | DRAM rd MB | DRAM wr MB | DRAM % | L2 % | L2 hit % | SM % | issue % | warps act % | regs | warp-instr M |\n")
    f.write("|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for k in kernels:
        f.write("| `%s` | %.1f | %.1f | %.1f | %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %d | %.1f |\n" % (
            k["kernel"].split("(")[0], k.get("time_us", 0), k.get("dram_read_bytes", 0) / 1e6, k.get("dram_write_bytes", 0) / 1e6,
            k.get("dram_pct", 0), k.get("l2_pct", 0), k.get("l2_hit_pct", 0), k.get("sm_pct", 0), k.get("issue_active_pct", 0),
            k.get("warps_active_pct", 0), int(k.get("registers", 0)), k.get("warp_instructions", 0) / 1e6))
    if shares:
        f.write("\n## Launch list (mean over %d steps)\n\n| kernel | mean us | share of step |\n|---|---|---|\n" %
                max(v["launches"] for v in shares.values()))
        for k, v in sorted(shares.items(), key=lambda kv: -kv[1]["share"]):
            f.write("| `%s` | %.1f | %.1f%% |\n" % (k.split("(")[0], v["mean_us"], 100 * v["share"]))
    if bench:
        f.write("\n## bench.py line of the same build\n\n```json\n%s\n```\n" % json.dumps(bench, indent=1))
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep, "6"], capture_output=True, text=True).stdout
    f.write("\n## Stall reasons and hottest SASS per kernel\n\n```\n%s\n```\n" % hot[hot.find("##"):] if "##" in hot else "")
print("wrote profiles/%s_*" % tag)
