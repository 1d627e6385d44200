#!/usr/bin/env python3
"""Throughput of the primer enumeration kernel on the +-200 flanks of every candidate of one
chromosome of the benchmark genome: python tools/primer_bench.py [chromosome index]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from cropsr_b200 import engine, primers

k = int(sys.argv[1]) if len(sys.argv) > 1 else 1
engine.init(0)
tok = bench.synth_tokens("arabidopsis")[k]
g = engine.Genome()
g.add_token(tok)
g.commit()
r = g.scan(20)
for strand in "+-":
    ex = r.extras(0, strand, 200)
    n = len(ex["cut"])
    seg = np.zeros(n, np.uint32)
    for rep in range(3):
        t0 = time.perf_counter()
        out = primers.design_windows(g, seg, ex["flank_lo"], ex["flank_hi"])
        dt = time.perf_counter() - t0
    ok = out["status"] == 0
    print(f"strand {strand}: {n} windows in {dt * 1e3:.1f} ms wall (H2D of the windows and D2H of 25 B/window included) = "
          f"{n / dt / 1e6:.1f} M windows/s, {n * 2000 / dt / 1e9:.1f} G primers/s; {int(ok.sum())} designed, "
          f"{int(out['n_pairs'][ok].sum())} pairs, {int((out['n_pairs'][ok] > 0).sum())} windows with a pair")
r.free()
g.free()
