#!/usr/bin/env python
"""Aggregate scripts/ncu_by_line.py output by enclosing function (nearest preceding RTT_HD / __device__ / template
function header in the source file).  Usage: ncu_by_func.py <report> <kernel substring> [variant]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_by_line.py"), sys.argv[1], sys.argv[2],
                      sys.argv[3] if len(sys.argv) > 3 else "fast", "100000"], capture_output=True, text=True).stdout
heads = {}
def funcs_of(f):
    if f in heads: return heads[f]
    p = os.path.join(ROOT, "raytracetorch_b200", "csrc", f)
    res = []
    if os.path.exists(p):
        for n, l in enumerate(open(p).read().splitlines(), 1):
            m = re.match(r"^(?:template.*>\s*)?(?:RTT_HD|__device__|__global__|inline|static)[^;]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
            if m and not l.strip().startswith("//"):
                res.append((n, m.group(1)))
    heads[f] = res
    return res
agg = {}
tot = 0.0
for l in out.splitlines():
    m = re.match(r"\s*([\d.]+)%\s+smp\s+(\d+)\s+(\S+):(\d+)", l)
    if not m: continue
    pct, smp, f, ln = float(m.group(1)), int(m.group(2)), m.group(3), int(m.group(4))
    name = "?"
    for n, fn in funcs_of(f):
        if n <= ln: name = fn
        else: break
    k = f"{f}:{name}"
    a = agg.setdefault(k, [0.0, 0])
    a[0] += pct; a[1] += smp
    tot += pct
print(out.splitlines()[0] if out else "no output")
for k, (pct, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{pct:6.1f}%  smp {smp:7d}  {k}")
