#!/usr/bin/env python3
"""Short driver for ncu: pack the benchmark genome once, run a few scans.
usage: python tools/profile_scan.py [workload] [n_scans]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from cropsr_b200 import engine

workload = sys.argv[1] if len(sys.argv) > 1 else "arabidopsis"
n_scans = int(sys.argv[2]) if len(sys.argv) > 2 else 3
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
engine.init(0)
g = engine.Genome()
for t in bench.synth_tokens(workload):
    g.add_token(t)
g.commit()
for _ in range(n_scans):
    r = g.scan(20, flags)
    print(f"{workload}: {g.num_positions} positions, {r.n_plus + r.n_minus} candidates, scan {r.scan_ms():.4f} ms, pack {g.timing()['pack_ms']:.4f} ms")
    r.free()
g.free()
