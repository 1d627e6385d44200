"""CPU ORACLE for the opt-in side outputs (GC, poly-T, homopolymer, cut site, +-L flank,
feature annotation).  TEST INFRASTRUCTURE ONLY.

Parity status: UNPINNED BY THE REFERENCE.  The reference computes none of these at this
commit (-L and -g are parsed and ignored, /root/reference/CROPSR.py:41,375; the GC rule
exists only in dead code, /root/reference/cropsr_functions.py:174-179), so there is no
reference output to pin against; this module is the specification the CUDA kernels are
tested against, written independently with Python strings.
"""
import numpy as np

import cropsr_oracle as oracle


def extras_for_token(key, tok, flank=200):
    """-> list of dicts, one per candidate with a full 30-base window, in reference order
    ('+' by ascending t, then '-'), each with t, strand, gc, flags, run, cut, flank_lo, flank_hi."""
    out = []
    for c in oracle.candidates_for_token(key, tok, 20):
        start, end, _, _, long_, _, strand = c
        t = end if strand == "+" else end - 3
        row = {"t": t, "strand": strand, "full": len(long_) == 30}
        if len(long_) == 30:
            scored = oracle.scored_bytes(long_)                 # upper-cased, U -> T
            proto = bytes(scored[5:25]).decode("latin-1")
            gc = proto.count("G") + proto.count("C")
            run, best, prev = 0, 0, None
            for ch in proto:
                if ch in "ACGT" and ch == prev:
                    run += 1
                elif ch in "ACGT":
                    run = 1
                else:
                    run = 0
                prev = ch if ch in "ACGT" else None
                best = max(best, run)
            flags = (1 if "TTTT" in proto else 0) | (2 if best >= 5 else 0) | (4 if gc < 10 else 0) | \
                    (8 if any(ch not in "ACGT" for ch in proto) else 0)
            row.update(gc=gc, flags=flags, run=best)
        cut = end - 3                                           # CROPSR.py:155-158
        row.update(cut=cut, flank_lo=max(cut - flank, 0), flank_hi=min(cut + flank, len(tok)))
        out.append(row)
    return out


def annotate(cuts, start, end):
    """Index of the containing interval with the largest index (sorted by start), else -1."""
    res = []
    for cut in cuts:
        hit = -1
        for j in range(len(start)):
            if start[j] <= cut <= end[j]:
                hit = j
        res.append(hit)
    return np.array(res, dtype=np.int32)


def other_runs(token, min_len=1):
    """Gap table: maximal runs of bytes that are not A C G T a c g t, at least min_len long.
    -> [(start, length)] ascending (UNPINNED BY THE REFERENCE, like everything in this module)."""
    import re
    return [(m.start(), m.end() - m.start()) for m in re.finditer(r"[^ACGTacgt]+", token) if m.end() - m.start() >= min_len]
