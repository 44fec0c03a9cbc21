#!/usr/bin/env python
"""Instructions per ray by SASS instruction and by source line, from the CSV of scripts/gpu_profile_sasscsv.sh.
   sass_by_line.py <sasscsv.csv.gz> <library .so> <mangled kernel substring> <rays> [annotated listing out]
Joins ncu's per-instruction executed counts with `nvdisasm -g` line info of the library that ran."""
import collections, csv, gzip, os, re, subprocess, sys, tempfile
rep, lib, kern, n = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
out = sys.argv[5] if len(sys.argv) > 5 else None
rows = list(csv.reader(gzip.open(rep, "rt")))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}
data = []
for r in rows[2:]:
    try:
        data.append((int(r[0], 16), r[1].strip(), float(r[ix["Instructions Executed"]] or 0), float(r[ix["# Samples"]] or 0)))
    except (ValueError, IndexError):
        pass
base = data[0][0]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if "kernels_fast" in f][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
addr2, cur, infn = {}, None, False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        infn = kern in ln; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and cur: addr2[int(m.group(1), 16)] = cur
CTL = ("BRA", "BSSY", "BSYNC", "BREAK", "BRX", "YIELD", "WARPSYNC", "NOP", "VOTE", "CALL", "RET")
agg, ctl, mn = collections.Counter(), collections.Counter(), collections.Counter()
lines = []
for a, s, ie, sm in data:
    off = a - base; c = ie * 32 / n; src = addr2.get(off, "?")
    agg[src] += c
    t = s.split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    mn[op] += c
    if op in CTL: ctl[src] += c
    lines.append(f"{off:06x} {c:7.2f} {int(sm):5d} {src:28s} {s}\n")
if out: open(out, "w").writelines(lines)
tot = sum(agg.values())
print(f"thread instructions per ray {tot:.1f} (control flow {sum(ctl.values()):.1f})")
print("by mnemonic:", ", ".join(f"{m} {c:.1f}" for m, c in mn.most_common(16)))
for s, c in agg.most_common(40): print(f"  {s:30s} {c:7.1f} {100 * c / tot:5.1f}%   control {ctl[s]:5.1f}")
