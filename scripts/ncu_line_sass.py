#!/usr/bin/env python
"""SASS instructions attributed to given source lines with executed counts.
   ncu_line_sass.py <report> <kernel substring> <file:line>[,<file:line>...] [variant]"""
import csv, os, re, subprocess, sys, tempfile
rep, kern, wanted = sys.argv[1], sys.argv[2], sys.argv[3].split(",")
variant = sys.argv[4] if len(sys.argv) > 4 else "fast"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "raytracetorch_b200", "librtt_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f"kernels_{variant}" in f][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
addr2, cur, infn = {}, None, False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        infn = kern in ln; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and cur: addr2[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, base_addr = None, None
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if not hdr or len(r) != len(hdr): continue
    try: a = int(r[0], 16) if not r[0].isdigit() else int(r[0])
    except ValueError: continue
    if base_addr is None: base_addr = a
    a -= base_addr
    if a in addr2 and addr2[a][0] in wanted:
        ie = float(r[hdr.index("Instructions Executed")] or 0)
        print(f"{a:06x} {addr2[a][0]:24s} {ie:12.0f}  {addr2[a][1][:90]}")
