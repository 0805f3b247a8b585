#!/usr/bin/env python
"""Top source lines of an `ncu --import-source on` report: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv`
then `python tools/ncu_hot_lines.py f.csv [n_envs] [top]`.  Aggregates the per-line rows of the FIRST kernel in the file."""
import csv
import os
import sys

def num(x):
    try:
        return float(x)
    except (TypeError, ValueError):
        return 0.0


rows = list(csv.reader(open(sys.argv[1])))
n_envs = float(sys.argv[2]) if len(sys.argv) > 2 else 4096.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lines = {}
fpath, kernel, first_kernel, hdr = None, None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        kernel = r[1]
        first_kernel = first_kernel or kernel
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if kernel != first_kernel or hdr is None or not r[0].strip().isdigit():
        continue
    d = dict(zip(hdr[4:], r[4:]))
    key = (fpath, int(r[0]))
    ent = lines.setdefault(key, dict(src=r[1].strip(), inst=0.0, samples=0.0))
    ent["inst"] += num(d.get("Instructions Executed"))
    ent["samples"] += num(d.get("# Samples"))
tot_i = sum(e["inst"] for e in lines.values()); tot_s = sum(e["samples"] for e in lines.values()) or 1.0
print(f"kernel: {first_kernel}\nwarp-inst/env: {tot_i / n_envs:.0f}, samples: {tot_s:.0f}")
for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    print(f"{f}:{ln:<5d} inst/env {e['inst'] / n_envs:7.1f}  samples {100 * e['samples'] / tot_s:5.1f}%  {e['src'][:110]}")
