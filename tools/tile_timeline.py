#!/usr/bin/env python3
"""Per-tile timeline of the scan kernel (debug): when each tile's counts were
ready, published, its look-back finished, and how long the workers waited."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cropsr_b200 import engine, _native

workload = sys.argv[1] if len(sys.argv) > 1 else "arabidopsis"
engine.init(0)
g = engine.Genome()
for t in bench.synth_tokens(workload):
    g.add_token(t)
g.commit()
n_tiles = (g.num_positions // engine.TILE) + 64
buf = torch.zeros(8 * n_tiles, dtype=torch.int64, device="cuda")
_native.lib.crp_debug_set_tile_times.argtypes = [C.c_void_p]
for rep in range(3):
    buf.zero_()
    torch.cuda.synchronize()
    _native.check(_native.lib.crp_debug_set_tile_times(buf.data_ptr()))
    r = g.scan(20)
    ms = r.scan_ms()
    r.free()
_native.check(_native.lib.crp_debug_set_tile_times(None))
T = buf.cpu().numpy().reshape(-1, 8)
T = T[T[:, 0] > 0]
t0 = T[:, :7][T[:, :7] > 0].min()
rel = (T[:, :7] - t0) / 1000.0
names = ["p1_done", "svc_ready", "svc_got_tot", "lb_done", "p2_wait0", "p2_wait1", "p2_done"]
print(f"scan {ms:.3f} ms, {len(T)} tiles, grid-stride view (us):")
for i in list(range(0, 6)) + list(range(440, 450)) + list(range(2000, 2004)) + list(range(len(T) - 4, len(T))):
    if i < len(T):
        print(i, " ".join(f"{n}={rel[i, k]:8.1f}" for k, n in enumerate(names)))
d = lambda a, b: rel[:, names.index(b)] - rel[:, names.index(a)]
w = T[:, 7] >> 32
pl = T[:, 7] & 0xFFFFFFFF
print("look-back windows: mean %.2f max %d ; max polls per lane: mean %.2f p90 %.0f max %d" % (w.mean(), w.max(), pl.mean(), np.percentile(pl, 90), pl.max()))
for a, b in (("p1_done", "svc_got_tot"), ("svc_got_tot", "lb_done"), ("p1_done", "lb_done"), ("p2_wait0", "p2_wait1"),
             ("p2_wait1", "p2_done"), ("p1_done", "p2_wait0")):
    x = d(a, b)
    print(f"{a:>12} -> {b:<12} mean {x.mean():8.2f}  p50 {np.median(x):8.2f}  p90 {np.percentile(x, 90):8.2f}  max {x.max():8.2f}")
