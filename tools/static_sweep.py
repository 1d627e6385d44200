import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import numpy as np
import workloads as W
from cropsr_b200 import engine
engine.init(0)
for wl in sys.argv[1:]:
    g = engine.Genome()
    for k in range(len(W.lengths(wl))):
        g.add_token(W.token(wl, k))
    g.commit()
    for e in ("0", "2", "4", "6", "8"):
        os.environ["CRP_STATIC_EIGHTHS"] = e
        ms = []
        for i in range(8):
            engine.flush_l2()
            r = g.scan(20)
            if i >= 3: ms.append(r.timing_detail()["kernel_ms"])
            r.free()
        print(wl, "static_eighths", e, "ms %.4f" % np.mean(ms), "min %.4f" % np.min(ms), flush=True)
    g.free()
