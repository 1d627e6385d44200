#!/usr/bin/env python3
"""Shared-memory wavefronts per SASS instruction of an .ncu-rep (top N), with the ideal count.
  python tools/ncu_smem.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
iw, ii, ie, isrc = (hdr.index(n) for n in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "Instructions Executed", "Source"))
data = [(int(r[iw]), int(r[ii]), int(r[ie]), r[isrc].strip()) for r in rows
        if len(r) == len(hdr) and r[iw].isdigit() and int(r[iw]) > 0]
tot = sum(d[0] for d in data)
print(f"shared wavefronts {tot}, ideal {sum(d[1] for d in data)}")
for w, i, e, src in sorted(data, reverse=True)[:top]:
    print(f"{100.0 * w / tot:5.1f}%  {w:9d} wavefronts  ideal {i:9d}  {w / max(e, 1):5.2f}/inst  x{e:8d}  {src[:70]}")
