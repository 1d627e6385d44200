"""Which summation order OpenBLAS gives each row of the reference's two
``np.matmul(matrix, weights)`` calls (/root/reference/CROPSR.py:305,311).

The reference scores every emitted slice with one matmul per weight vector, and
OpenBLAS 0.3.30's dgemv_t sums a row differently depending on where the row
sits in the call: rows are split into contiguous per-thread ranges, each range
is walked four rows at a time (CANONICAL lane order), a remainder of 2 or 3 rows
sends its first two through a 2-lane kernel (PAIR), and a one-row call is a ddot
(SINGLE).  The GPU scan computes CANONICAL for every candidate; the handful of
other rows are re-evaluated by ``crp_rescore``.  Host-side index logic only —
no floating point happens here.  
"""
import numpy as np

CANONICAL, PAIR, SINGLE = 0, 1, 2
FIRST_COLS, SECOND_COLS = 120, 464
MULTITHREAD_MIN_ELEMS = 460800     # gemv is threaded when rows*cols >= this


def thread_ranges(n_rows, n_cols, threads):
    if threads <= 1 or n_rows * n_cols < MULTITHREAD_MIN_ELEMS:
        return [(0, n_rows)]
    out, left, pos, used = [], n_rows, 0, 0
    while left > 0:
        width = (left + threads - used - 1) // (threads - used)
        width = min(max(width, 4), left)
        out.append((pos, pos + width))
        pos += width
        left -= width
        used += 1
    return out


def noncanonical_rows(n_rows, n_cols, threads=1):
    """{row index: class} for the rows of an (n_rows, n_cols) matmul that are
    not summed in the canonical order."""
    if n_rows == 1:
        return {0: SINGLE}
    out = {}
    for a, b in thread_ranges(n_rows, n_cols, threads):
        width = b - a
        if width % 4 in (2, 3):
            base = a + (width // 4) * 4
            out[base] = PAIR
            out[base + 1] = PAIR
    return out


def slice_classes(n_rows, threads=1):
    """{row index: first_class | second_class << 4} for rows of an emitted
    slice of n_rows whose x must be re-evaluated."""
    first = noncanonical_rows(n_rows, FIRST_COLS, threads)
    second = noncanonical_rows(n_rows, SECOND_COLS, threads)
    return {i: first.get(i, CANONICAL) | (second.get(i, CANONICAL) << 4)
            for i in sorted(set(first) | set(second))}
