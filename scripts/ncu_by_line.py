#!/usr/bin/env python
"""Aggregate an ncu source page (SASS level) by CUDA source line.

    python scripts/ncu_by_line.py <report.ncu-rep> <kernel mangled-name substring> [variant cubin] [top] [--by-samples]

The substring must select ONE function of the cubin: for a templated kernel pass the mangled instantiation
(e.g. k_trace_seq_bwd_fastILi4ELb0), otherwise the address maps of several instantiations overwrite each other.
--by-samples orders the lines by warp stall samples instead of executed instructions (where warps WAIT, e.g. the
compaction pass of the adjoint that executed 3 % of the instructions and held 24 % of the samples).

Joins `ncu --page source --csv` (per-SASS-address executed instructions and stall samples) with the
`//## File ..., line N` annotations of `nvdisasm -g` on the cubin extracted from librtt_b200.so
(compiled with -lineinfo).  Prints the source lines that execute the most instructions.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

BY_SAMPLES = "--by-samples" in sys.argv
sys.argv = [a for a in sys.argv if a != "--by-samples"]
rep, kern = sys.argv[1], sys.argv[2]
variant = sys.argv[3] if len(sys.argv) > 3 else "fast"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "raytracetorch_b200", "librtt_b200.so")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f"kernels_{variant}" in f][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout

# address -> (file, line) for the wanted function
addr2line, cur, infn = {}, None, False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        infn = kern in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and cur:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg, tot, hdr, active, taken, base_addr = {}, 0.0, None, False, False, None
_m = re.search(r"k_[a-z_]+?(?=_fast|_exact|$)", kern)
base = _m.group(0) if _m else "k_trace"
for r in rows:
    if r and r[0] == "Kernel Name":
        active = (base in r[1]) and not taken
        taken = taken or active
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if not (active and hdr) or len(r) != len(hdr):
        continue
    try:
        a = int(r[0], 16) if not r[0].isdigit() else int(r[0])
    except ValueError:
        continue
    ie = float(r[hdr.index("Instructions Executed")] or 0)
    smp = float(r[hdr.index("# Samples")] or 0)
    if base_addr is None:
        base_addr = a
    a -= base_addr
    key = addr2line[a][0] if a in addr2line else ("?", 0)
    d = agg.setdefault(key, [0.0, 0.0])
    d[0] += ie
    d[1] += smp
    tot += ie
srcs = {}
print(f"total warp instructions {tot:.3e}")
for key, (ie, smp) in sorted(agg.items(), key=lambda kv: -kv[1][1 if BY_SAMPLES else 0])[:top]:
    f, ln = key
    if f not in srcs:
        p = os.path.join(ROOT, "raytracetorch_b200", "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ""
    print(f"{ie / tot * 100:5.1f}%  smp {int(smp):6d}  {f}:{ln:<4d} {text}")
