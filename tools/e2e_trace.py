#!/usr/bin/env python3
"""Host-side stage trace of the pipelined call (CRP_TRACE=1 python tools/e2e_trace.py [workload])."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from cropsr_b200 import engine
workload = sys.argv[1] if len(sys.argv) > 1 else "arabidopsis"
engine.init(0)
toks = bench.synth_tokens(workload)
pinned = []
for t in toks:
    b = engine.PinnedBuffer(len(t)); b.array[:] = t; pinned.append(b)
segs = [(k, b.array, 0, None) for k, b in enumerate(pinned)]
arena = None
for rep in range(4):
    t0 = time.perf_counter()
    arena, n_plus, n_minus, ms = engine.scan_segments(segs, 20, arena=arena)
    print(f"rep {rep}: {1e3 * (time.perf_counter() - t0):.3f} ms wall, device scan {ms:.3f} ms, {int(n_plus.sum() + n_minus.sum())} rows", file=sys.stderr)
