"""Primer enumeration on the device for many fragments at once (opt-in side output).

Host-side mirror of the part of the reference's primer designer that needs no aligner
(/root/reference/prmrdsgn2.py): ``get_primers`` (:115-124) on a fragment and on its reverse
complement (:104-112), the ``Primer`` GC % / Tm (:76-95), ``filter_primers`` (:127-137) and the Tm
pairing of ``main()`` (:260-266).  The keyword arguments carry the reference's CLI flags
(-e -s -l -m -x -M -X -D, :26-54) with the same defaults.  The fragments are windows of a genome
that is already packed in HBM -- typically the +-L flank of a candidate's cut site, the
``flank_lo`` / ``flank_hi`` of ``ScanResult.extras`` -- and the work is done by ``k_primers``
through ``crp_primer_windows``; there is no CPU path.  bowtie2 alignment of the pairs
(:139-160) is out of scope.  CROPSR.py never calls prmrdsgn2, so nothing here enters the CSV.
"""
import ctypes as C

import numpy as np

from . import _native as N
from ._native import check, lib

DEFAULTS = dict(e=100, s=20, l=30, m=50.0, x=65.0, M=35.0, X=65.0, D=0.5)   # prmrdsgn2.py:26-54


def design_windows(genome, segment, lo, hi, **params):
    """Primers of the windows [lo[i], hi[i]) (token positions) of segment[i] of a committed Genome.

    -> dict(n_fwd, n_rev: uint32[n]; n_pairs: uint64[n]; first: uint16[n, 4] = forward (start,
    length), reverse (start on the reverse complement, length) of the first Tm-compatible pair in the
    reference's order, 0xFFFF if none; status: uint8[n], 1 where the window is shorter than e + l)."""
    unknown = set(params) - set(DEFAULTS)
    if unknown:
        raise TypeError(f"unknown primer parameter(s): {sorted(unknown)}")
    p = dict(DEFAULTS, **params)
    prm = N.PrimerParams(int(p["e"]), int(p["s"]), int(p["l"]), float(p["m"]), float(p["x"]), float(p["M"]),
                         float(p["X"]), float(p["D"]))
    segment = np.ascontiguousarray(segment, dtype=np.uint32)
    lo = np.ascontiguousarray(lo, dtype=np.uint32)
    hi = np.ascontiguousarray(hi, dtype=np.uint32)
    n = len(lo)
    if len(segment) != n or len(hi) != n:
        raise ValueError("segment, lo and hi must have the same length")
    out = {"n_fwd": np.zeros(n, np.uint32), "n_rev": np.zeros(n, np.uint32), "n_pairs": np.zeros(n, np.uint64),
           "first": np.full((n, 4), 0xFFFF, np.uint16), "status": np.zeros(n, np.uint8)}
    if n:
        check(lib.crp_primer_windows(genome._h, n, segment.ctypes.data, lo.ctypes.data, hi.ctypes.data, C.byref(prm),
                                     out["n_fwd"].ctypes.data, out["n_rev"].ctypes.data, out["n_pairs"].ctypes.data,
                                     out["first"].ctypes.data, out["status"].ctypes.data))
    return out
