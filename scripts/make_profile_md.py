#!/usr/bin/env python
"""Assemble profiles/<out>.md from the on-box ncu summaries (scripts/gpu_profile_any.sh / gpu_profile_tile.sh write
sum_/func_/samples_/raw_ text files into gpurun_out/), and update profiles/ncu_counters.json + profiles/traffic.json.

    python scripts/make_profile_md.py <out name> <title> <name:tag:kernel:workload[:rays]> ...
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
out_name, title, items = sys.argv[1], sys.argv[2], sys.argv[3:]
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic")


def read(path):
    try:
        return open(path).read()
    except OSError:
        return ""


def to_bytes(val, unit_line):
    v = float(val)
    u = unit_line.lower()
    return v * (1e9 if "gbyte" in u else 1e6 if "mbyte" in u else 1e3 if "kbyte" in u else 1.0)


counters_path = os.path.join(ROOT, "profiles", "ncu_counters.json")
traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
counters = json.loads(read(counters_path) or "{}")
traffic = json.loads(read(traffic_path) or "{}")
md = [f"# {title}", "",
      "Every capture: `ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1` of a `bench.py` "
      "command that had just exited 0 without ncu on the same box (scripts/gpu_profile_any.sh / gpu_profile_tile.sh), "
      "summarised on the box.  Durations are cold-cache, serialised launches: compare shares, not absolutes.", ""]
for it in items:
    parts = it.split(":")
    name, tag, kernel, wl = parts[:4]
    rays = int(float(parts[4])) if len(parts) > 4 else 20_000_000
    summ = read(os.path.join(G, f"sum_{name}_{tag}.txt"))
    raw = read(os.path.join(G, f"raw_{name}_{tag}.txt"))
    func = read(os.path.join(G, f"func_{name}_{tag}.txt"))
    samp = read(os.path.join(G, f"samples_{name}_{tag}.txt"))
    md += [f"## {name} — `{kernel}` on {wl}, {rays:.0e} rays", "", "```"]
    vals = {}
    for ln in summ.splitlines():
        m = re.match(r"\s+(\S+)\s+(\S+)\s*(.*)$", ln)
        if m and m.group(1) in KEEP:
            md.append(ln.rstrip())
            vals[m.group(1)] = (m.group(2), m.group(3))
        elif ln.startswith("====="):
            md.append(ln.rstrip())
    md.append("```")
    stalls = []
    for ln in raw.splitlines():
        m = re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+)_not_issued (\d+)", ln)
        if m:
            stalls.append((int(m.group(2)), m.group(1)))
    stalls.sort(reverse=True)
    tot = sum(s for s, _ in stalls) or 1
    md += ["", "Warp stall samples (not issued): " + ", ".join(f"{n} {100 * s / tot:.0f} %" for s, n in stalls[:7]), ""]
    md += ["Executed instructions by function:", "```"] + func.splitlines()[:16] + ["```", ""]
    md += ["Source lines holding the most stall samples:", "```"] + samp.splitlines()[1:13] + ["```", ""]
    key = f"{kernel}:{wl}"
    try:
        counters[key] = dict(fma_pipe_pct=float(vals["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"][0]),
                             issue_active_pct=float(vals["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
                             warps_active_pct=float(vals["sm__warps_active.avg.pct_of_peak_sustained_active"][0]),
                             registers=int(float(vals["launch__registers_per_thread"][0])),
                             source=f"profiles/{out_name}.md ({name}), committed ncu --set full capture; not measured in the bench run")
        rd = to_bytes(*vals["dram__bytes_read.sum"])
        wr = to_bytes(*vals["dram__bytes_write.sum"])
        traffic[key] = dict(rays=rays, dram_bytes=int(rd + wr),
                            source=f"profiles/{out_name}.md ({name}): dram__bytes_read.sum + dram__bytes_write.sum of one "
                                   f"launch of {rays:.0e} rays ({(rd + wr) / rays:.1f} B/ray)")
    except (KeyError, ValueError) as exc:
        md.append(f"(counters not parsed: {exc})")
open(os.path.join(ROOT, "profiles", out_name + ".md"), "w").write("\n".join(md) + "\n")
json.dump(counters, open(counters_path, "w"), indent=1)
json.dump(traffic, open(traffic_path, "w"), indent=1)
print("wrote", out_name)
