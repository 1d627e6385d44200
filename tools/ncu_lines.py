#!/usr/bin/env python3
"""Per-source-line instruction / stall summary of an .ncu-rep (needs -lineinfo).

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]

Reads `ncu --page source --print-source cuda,sass --csv` and prints, for the lines
of the profiled kernel that execute the most warp instructions: share of all
executed instructions, average active threads, stall samples and dominant stalls.
"""
import csv
import subprocess
import sys


def num(v):
    return int(v) if v.isdigit() else 0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    lines = []
    for r in rows:
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "" or not r[hdr.index("Instructions Executed")].isdigit():
            continue
        lines.append(r)
    ie = hdr.index("Instructions Executed")
    te = hdr.index("Thread Instructions Executed")
    smp = hdr.index("# Samples")
    stall0 = hdr.index("stall_barrier")
    stall1 = hdr.index("stall_wait") + 1
    names = hdr[stall0:stall1]
    tot_i = sum(int(r[ie]) for r in lines)
    tot_s = sum(num(r[smp]) for r in lines)
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    lines.sort(key=lambda r: -(num(r[smp]) if len(sys.argv) > 3 else num(r[ie])))
    for r in lines[:top]:
        i, t, s = num(r[ie]), num(r[te]), num(r[smp])
        st = sorted(((int(v), n) for v, n in zip(r[stall0:stall1], names) if v.isdigit() and int(v)), reverse=True)[:3]
        print(f"{100.0 * i / tot_i:5.1f}% inst  {100.0 * s / max(tot_s, 1):5.1f}% smp  thr {t / max(i, 1):4.1f}  L{r[0]:>4}: "
              f"{r[1].strip()[:90]}  {[f'{n[6:]}={v}' for v, n in st]}")


if __name__ == "__main__":
    main()
