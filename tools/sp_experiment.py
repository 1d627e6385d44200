#!/usr/bin/env python3
"""Debug timing of the single-pass kernel: normal launches, then launches that reuse the
prefixes of the previous launch (no look-back) -- the cost of the look-back is the difference."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from cropsr_b200 import engine

workload = sys.argv[1] if len(sys.argv) > 1 else "arabidopsis"
engine.init(0)
g = engine.Genome()
for t in bench.synth_tokens(workload):
    g.add_token(t)
g.commit()
for dbg in ("0", "0", "0", "0x10000000", "0x10000000", "0x10000000", "0x30000000", "0x20000000", "0"):
    os.environ["CRP_SP_DEBUG"] = dbg
    r = g.scan(20, 0)
    print(f"debug {dbg}: {r.n_plus + r.n_minus} candidates, scan {r.scan_ms():.4f} ms")
    r.free()
g.free()
