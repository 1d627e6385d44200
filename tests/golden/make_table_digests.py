#!/usr/bin/env python3
"""Digests of the unique-candidate table (pos, packed 30-mer, fp64 x per candidate, both strands,
every record) of the synthetic BASELINE configs, computed by the vectorised CPU oracle
(oracle/vector_oracle.py, pinned to the reference through the literal port).  The GPU parity
tests recompute the same digest from the device's streams (SURVEY.md 8d parity protocol:
"the spec/vectorised oracle's digest of the unique candidate table" for the configs the
reference itself cannot finish).

usage: python tests/golden/make_table_digests.py [workload ...]     -> tests/golden/table_digests.json
       (arabidopsis ~1 min, sorghum ~5 min, maize ~20 min, sugarcane ~2 h of one core)
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
OUT = os.path.join(HERE, "table_digests.json")


def main():
    import numpy as np
    import vector_oracle as vo
    import workloads as W
    names = sys.argv[1:] or ["arabidopsis"]
    table = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        t0 = time.time()
        d = vo.TableDigest()
        n_plus = n_minus = 0
        for k in range(len(W.lengths(name))):
            tab = vo.token_table(W.token(name, k), 20)
            d.add_token(k, tab["+"], tab["-"])
            n_plus += len(tab["+"][0])
            n_minus += len(tab["-"][0])
        table[name] = {"sha256": d.hexdigest(), "tokens": d.tokens, "candidates": d.candidates, "plus": n_plus,
                       "minus": n_minus, "positions": sum(W.token_lengths(name)), "numpy": np.__version__}
        print(name, table[name], f"{time.time() - t0:.0f} s", flush=True)
        json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
