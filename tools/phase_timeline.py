#!/usr/bin/env python3
"""Phase timeline of the scan kernel (debug): per CTA, when the table load, count phase,
grid barrier and emit phase ended.  usage: python tools/phase_timeline.py [workload] [flags]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # device buffer only
import bench
from cropsr_b200 import engine, _native

workload = sys.argv[1] if len(sys.argv) > 1 else "arabidopsis"
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
engine.init(0)
g = engine.Genome()
for t in bench.synth_tokens(workload):
    g.add_token(t)
g.commit()
buf = torch.zeros(8 * 1024, dtype=torch.int64, device="cuda")
_native.lib.crp_debug_set_times.argtypes = [C.c_void_p]
for rep in range(3):
    buf.zero_()
    torch.cuda.synchronize()
    _native.check(_native.lib.crp_debug_set_times(buf.data_ptr()))
    r = g.scan(20, flags)
    ms = r.scan_ms()
    r.free()
_native.check(_native.lib.crp_debug_set_times(None))
T = buf.cpu().numpy().reshape(-1, 8)
T = T[T[:, 0] > 0]
t0 = T[:, 0].min()
rel = (T - t0) / 1000.0
names = ["start", "count_begin", "count_end", "grid_sync_end", "emit_begin", "emit_end", "slot6", "slot7"]
print(f"scan {ms * 1e3:.1f} us (events), {len(T)} CTAs; microseconds since the first CTA started")
for k, n in enumerate(names):
    c = rel[:, k]
    if not np.isfinite(c).all() or abs(c).max() > 1e7:
        continue
    print(f"{n:>14}: min {c.min():7.1f}  p50 {np.median(c):7.1f}  p90 {np.percentile(c, 90):7.1f}  max {c.max():7.1f}")
d = rel[:, 5] - rel[:, 4]
print(f"emit duration per CTA: min {d.min():.1f} p50 {np.median(d):.1f} max {d.max():.1f}")

if os.environ.get("CRP_WS_WAITSTATS"):
    # slots 1, 6, 7 hold cycle counts (front-end warp 0 waiting for records, loader waiting for a free
    # stage, first body warp waiting for lists), accumulated since the buffer was zeroed
    for k, n in ((1, "fe wait rec_full"), (6, "loader wait empty"), (7, "body wait list_full")):
        c = T[:, k].astype(float) / 1.9e3
        print(f"{n:>22}: p50 {np.median(c):7.1f} us  max {c.max():7.1f} us (at 1.9 GHz)")
