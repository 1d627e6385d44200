"""VECTORISED CPU ORACLE of the unique-candidate table (numpy), for inputs the literal port in
cropsr_oracle.py is too slow for (the 135 Mbp - 10 Gbp configs).

TEST INFRASTRUCTURE ONLY -- same rules as cropsr_oracle.py: tests/, smoke() and bench.py's CPU
legs may import it, the product never does.

It is written from the spec, not from the device code:
  * PAM positions: regex `(?=.GG)` / `(?=CC.)` on the token, case-sensitive, with the bounds of
    /root/reference/CROPSR.py:419 and :430                                   -> `pam_positions`
  * the scored 30-mer of a hit: the per-byte effect of the reference's own replace chains
    (`get_gRNA_sequence` :124-129, `get_reverse_complement` :116-121) followed by
    `.replace('U','T').upper()` (:458).  The two 256-entry byte tables are DERIVED by pushing
    every byte through cropsr_oracle.grna / reverse_complement, so nothing is restated twice
                                                                              -> `byte_tables`
  * x = -(((A+B)+0.59763615)+(-0.2026259)) with A, B summed in OpenBLAS' canonical lane order
    (four sequential accumulators by column mod 4, (p0+p2)+(p1+p3)) -> `canonical_x`
  * the packed word of include/cropsr_b200.h (planar codes of the scoring bases + flags)

Parity status: PINNED through cropsr_oracle.py -- tests/test_oracle_golden.py checks this module
against the literal port (positions, 30-mers, x bit for bit) on every golden fixture and on
random FASTAs; the literal port is pinned to the unmodified reference's CSVs.
"""
import hashlib

import numpy as np

import cropsr_oracle as lit
from rs1_table import W1, W2, INTERCEPT, LOW_GC

NO_SCORE = 255
_CODE = {ord("A"): 0, ord("T"): 1, ord("C"): 2, ord("G"): 3}          # CROPSR.py:300-302


def byte_tables():
    """plus[c] / minus[c]: code (0..3) the token byte c contributes to the scored 30-mer of a
    '+' / '-' hit, or NO_SCORE.  Derived from the literal string transforms."""
    plus = np.full(256, NO_SCORE, dtype=np.uint8)
    minus = np.full(256, NO_SCORE, dtype=np.uint8)
    for c in range(128):                       # tokens are ASCII (the reference's bytes(..., "ascii") raises otherwise)
        ch = chr(c)
        p = lit.grna(ch).replace("U", "T").upper()
        m = lit.grna(lit.reverse_complement(ch)).replace("U", "T").upper()
        if len(p) == 1 and ord(p) in _CODE:
            plus[c] = _CODE[ord(p)]
        if len(m) == 1 and ord(m) in _CODE:
            minus[c] = _CODE[ord(m)]
    return plus, minus


_PLUS, _MINUS = byte_tables()
_UPPER_ACGT = np.zeros(256, dtype=bool)
_UPPER_ACGT[[ord(c) for c in "ACGT"]] = True


def pam_positions(tok, guide_len=20):
    """tok: uint8 array of the token.  -> (t_plus, t_minus) int64, ascending (CROPSR.py:415-430)."""
    L = len(tok)
    if L < 3:
        return np.empty(0, np.int64), np.empty(0, np.int64)
    g, c = tok == ord("G"), tok == ord("C")
    plus = np.nonzero(g[1:L - 1] & g[2:L])[0]                      # t + 2 <= L - 1
    plus = plus[plus >= guide_len + 5]                             # p0 = t - l >= 5
    minus = np.nonzero(c[0:L - 2] & c[1:L - 1])[0]                 # a third byte exists
    minus = minus[(minus >= 2) & (minus <= L - guide_len + 7)]     # p0 = t + 3 >= 5, p1 = t + 3 + l <= L + 10
    return plus.astype(np.int64), minus.astype(np.int64)


def windows(tok, t, minus):
    """(n, 30) token bytes of the scored window of every hit in OUTPUT order (the order of the CSV's
    long_sequence), 0 where the window leaves the token; plus the truncated flag."""
    L = len(tok)
    q = np.arange(30, dtype=np.int64)
    idx = (t[:, None] - 2 + q[None, :]) if minus else (t[:, None] + 4 - q[None, :])
    inside = (idx >= 0) & (idx < L)
    w = np.where(inside, tok[np.clip(idx, 0, max(L - 1, 0))], 0).astype(np.uint8)
    # Python slices truncate at the token end (and a negative start cannot happen inside the bounds)
    truncated = ~inside.all(axis=1)
    return w, truncated


def canonical_x(code):
    """code: (n, 30) uint8 in {0..3, NO_SCORE}.  -> x float64, canonical lane order."""
    n = len(code)
    lanes1 = [np.zeros(n) for _ in range(4)]
    for p in range(30):
        for c in range(4):
            w = W1[4 * p + c]
            if w != 0.0:
                lanes1[c] = lanes1[c] + np.where(code[:, p] == c, w, 0.0)
    a = (lanes1[0] + lanes1[2]) + (lanes1[1] + lanes1[3])
    lanes2 = [np.zeros(n) for _ in range(4)]
    for p in range(29):
        for c1 in range(4):
            for c2 in range(4):                                     # column j = 16 p + 4 c1 + c2, lane j mod 4 = c2
                w = W2[16 * p + 4 * c1 + c2]
                if w != 0.0:
                    lanes2[c2] = lanes2[c2] + np.where((code[:, p] == c1) & (code[:, p + 1] == c2), w, 0.0)
    b = (lanes2[0] + lanes2[2]) + (lanes2[1] + lanes2[3])
    return (a + b + INTERCEPT + LOW_GC) * -1


def packed_words(code, window_bytes, truncated):
    """the uint64 of include/cropsr_b200.h: planar code bits of the scoring bases, bit 30 irregular,
    bit 31 truncated, bit 62 unscored"""
    scoring = code != NO_SCORE
    c = np.where(scoring, code, 0).astype(np.uint64)
    sh = np.arange(30, dtype=np.uint64)
    lo = ((c & np.uint64(1)) << sh).sum(axis=1, dtype=np.uint64)
    hi = ((c >> np.uint64(1)) << sh).sum(axis=1, dtype=np.uint64)
    irregular = ~_UPPER_ACGT[window_bytes].all(axis=1)
    out = lo | (hi << np.uint64(32))
    out |= np.where(irregular, np.uint64(1 << 30), np.uint64(0))
    out |= np.where(truncated, np.uint64(1 << 31), np.uint64(0))
    out |= np.where(~scoring.all(axis=1), np.uint64(1 << 62), np.uint64(0))
    return out


def token_table(tok, guide_len=20, chunk=2_000_000):
    """Unique candidates of one token in reference order.
    -> {'+': (pos u32, packed u64, x f64), '-': (...)}; packed / x are None unless guide_len == 20."""
    tok = np.frombuffer(tok, dtype=np.uint8) if not isinstance(tok, np.ndarray) else tok
    out = {}
    for strand, t in zip("+-", pam_positions(tok, guide_len)):
        if guide_len != 20:
            out[strand] = (t.astype(np.uint32), None, None)
            continue
        packed = np.empty(len(t), np.uint64)
        x = np.empty(len(t), np.float64)
        for lo in range(0, len(t), chunk):
            w, trunc = windows(tok, t[lo:lo + chunk], strand == "-")
            code = (_MINUS if strand == "-" else _PLUS)[w]
            code[w == 0] = NO_SCORE
            packed[lo:lo + chunk] = packed_words(code, w, trunc)
            x[lo:lo + chunk] = canonical_x(code)
        out[strand] = (t.astype(np.uint32), packed, x)
    return out


class TableDigest:
    """sha256 over the unique-candidate table: per token a header (index, n_plus, n_minus), then the
    little-endian bytes of pos / packed / x of the '+' stream, then of the '-' stream."""

    def __init__(self):
        self.h = hashlib.sha256()
        self.tokens = 0
        self.candidates = 0

    def add_token(self, index, plus, minus):
        n_p, n_m = len(plus[0]), len(minus[0])
        self.h.update(b"T" + np.array([index, n_p, n_m], dtype="<u8").tobytes())
        for pos, packed, x in (plus, minus):
            self.h.update(np.ascontiguousarray(pos, dtype="<u4").tobytes())
            if packed is not None:
                self.h.update(np.ascontiguousarray(packed, dtype="<u8").tobytes())
                self.h.update(np.ascontiguousarray(x, dtype="<f8").tobytes())
        self.tokens += 1
        self.candidates += n_p + n_m

    def hexdigest(self):
        return self.h.hexdigest()


def digest_of_tokens(tokens, guide_len=20):
    d = TableDigest()
    for k, tok in enumerate(tokens):
        tab = token_table(tok, guide_len)
        d.add_token(k, tab["+"], tab["-"])
    return d
