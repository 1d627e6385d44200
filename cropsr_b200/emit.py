"""Row emission: the reference's chunked, cumulative CSV writing, bug for bug.

Host-side mirror of /root/reference/CROPSR.py:386-405 (header), :442-474
(cumulative list, 1,000,000-row chunk plan, id indexing, row tuples),
:155-158 (cut site) and :316-318 (ids).  All arithmetic on scores' inputs was
done on the GPU (x per candidate, plus crp_rescore for the rows OpenBLAS sums
in another lane order); here x only goes through the same ``1/(1+np.exp(x))``
numpy expression the reference evaluates (CROPSR.py:313).
"""
import csv

import numpy as np

from . import blas_order

HEADER = ["crispr_id", "crispr_sys", "sequence", "long_sequence", "chromosome", "start_pos",
          "end_pos", "cutsite", "strand", "on_site_score", "features", "status"]
CHUNK_ROWS = 1000000

alphanum = np.array(list("ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"), dtype="|U1")


def get_id(num_to_gen):
    """CROPSR.py:316-318 -- same call on numpy's global legacy RNG, so a caller
    that seeds np.random gets the reference's ids."""
    return np.random.choice(alphanum, [num_to_gen, 7])


def legacy_id_bytes(num_to_gen):
    """The characters of get_id(num_to_gen) as (n, 7) ASCII bytes, drawn by the library's own MT19937
    loop (crp_legacy_ids, csrc/emit_csv.cpp) from numpy's global legacy generator: same values, same
    generator state afterwards, ~10x faster than numpy's bounded-integer path -- on a 120 Mbp genome
    the ids were half of the CLI's wall time."""
    import ctypes as C
    from ._native import lib, check
    kind, key, pos, has_gauss, cached = np.random.get_state()
    if kind != "MT19937":
        return id_bytes_of(get_id(num_to_gen))
    key = np.ascontiguousarray(key, dtype=np.uint32).copy()
    p = C.c_int32(int(pos))
    out = np.empty((int(num_to_gen), 7), dtype=np.uint8)
    check(lib.crp_legacy_ids(key.ctypes.data, C.byref(p), int(num_to_gen), out.ctypes.data if num_to_gen else None))
    np.random.set_state((kind, key, p.value, has_gauss, cached))
    return out


class IdStream:
    """get_id() for a known sequence of sizes, drawn ahead on a helper thread (crp_legacy_ids releases
    the GIL): the ids of the next emission are generated while the current one is formatted.  The
    generator state is taken from numpy when the stream is opened and handed back by close(), so the
    values and the state afterwards are those of calling get_id(size) for every size in turn; nothing
    else may draw from np.random in between."""

    def __init__(self, sizes):
        import queue
        import threading
        self.sizes = [int(n) for n in sizes]
        self.state = np.random.get_state()
        self.q = queue.Queue(maxsize=1)
        self.thread = None
        self.k = 0
        if self.state[0] == "MT19937" and self.sizes:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def _run(self):
        import ctypes as C
        from ._native import lib
        kind, key, pos, has_gauss, cached = self.state
        key = np.ascontiguousarray(key, dtype=np.uint32).copy()
        p = C.c_int32(int(pos))
        for n in self.sizes:
            out = np.empty((n, 7), dtype=np.uint8)
            rc = lib.crp_legacy_ids(key.ctypes.data, C.byref(p), n, out.ctypes.data if n else None)
            self.q.put(out if rc == 0 else None)
        self.state = (kind, key, p.value, has_gauss, cached)

    def next(self, size):
        if self.thread is None:
            return legacy_id_bytes(size)
        assert self.k < len(self.sizes) and self.sizes[self.k] == int(size), "IdStream: sizes out of sequence"
        self.k += 1
        out = self.q.get()
        if out is None:
            raise RuntimeError("crp_legacy_ids failed")
        return out

    def close(self):
        if self.thread is not None:
            while self.k < len(self.sizes):            # drain what was never asked for
                self.q.get()
                self.k += 1
            self.thread.join()
            np.random.set_state(self.state)
            self.thread = None


def apply_cutsite(start_pos, end_pos, crispr_sys):
    """CROPSR.py:155-158."""
    if crispr_sys == "cas9":
        cutsite = end_pos - 3
    return cutsite


def emission_slices(size, chunk_rows=None):
    """(start, count) of every slice of a cumulative list of `size` rows that
    the reference scores and writes (CROPSR.py:451-472): full 1e6-row slices,
    then -- because the start of the last slice is computed as count*counter --
    a last partial slice that starts at r*q instead of 1e6*q; when size is an
    exact multiple of 1e6 the final full slice is never written.
    chunk_rows: the reference's literal 1000000 (CROPSR.py:453); tests shrink it to walk the
    same plan on small inputs."""
    chunk = CHUNK_ROWS if chunk_rows is None else int(chunk_rows)
    q, r = divmod(size, chunk)
    if r:
        return [(chunk * j, chunk) for j in range(q)] + [(r * q, r)]
    return [(chunk * j, chunk) for j in range(max(q - 1, 0))]


# byte translation of the reference's replace chains (CROPSR.py:120,128)
def _table(pairs):
    t = bytearray(range(256))
    for a, b in pairs:
        t[ord(a)] = ord(b)
    return bytes(t)


PLUS_TABLE = _table([("A", "U"), ("C", "G"), ("G", "C"), ("T", "A"), ("Z", "G")])    # gRNA(x), then reversed
MINUS_TABLE = _table([("T", "U"), ("U", "A"), ("Z", "C")])                            # gRNA(revcomp(x)): forward


def guide_strings(token, t, minus, guide_len):
    """(short, long) CSV strings of the hit at regex position t (CROPSR.py:418-433),
    with Python's silent slice truncation at the token end."""
    if minus:
        p0 = t + 3
        p1 = p0 + guide_len
        return (token[p0:p1].translate(MINUS_TABLE).decode("ascii"),
                token[p0 - 5:p1 + 5].translate(MINUS_TABLE).decode("ascii"))
    p0 = t - guide_len
    return (token[p0:t].translate(PLUS_TABLE)[::-1].decode("ascii"),
            token[p0 - 5:t + 5].translate(PLUS_TABLE)[::-1].decode("ascii"))


class CandidateTable:
    """Unique candidates of all tokens seen so far, in reference order
    (per token: '+' by ascending t, then '-' by ascending t).  The columns live in arrays
    that are allocated once when the total is known (`capacity`: the scan has counted every
    token before the first row is written) and otherwise grow geometrically -- appending a token
    never copies the rows of the tokens before it (20,000 scaffolds: no O(k^2) concatenation)."""

    _COLS = (("tok", np.uint32), ("seg", np.uint32), ("minus", np.bool_), ("t", np.uint32), ("x", np.float64))

    def __init__(self, guide_len, capacity=0):
        self.guide_len = guide_len
        self.tokens = []      # bytes per token
        self.chroms = []      # CSV chromosome string per token (key[1:])
        self.n = 0
        self._cap = int(capacity)
        self._store = {name: np.empty(self._cap, dt) for name, dt in self._COLS}

    def __len__(self):
        return self.n

    tok = property(lambda self: self._store["tok"][:self.n])
    seg = property(lambda self: self._store["seg"][:self.n])
    minus = property(lambda self: self._store["minus"][:self.n])
    t = property(lambda self: self._store["t"][:self.n])
    x = property(lambda self: self._store["x"][:self.n])

    @property
    def token_len(self):
        """int64 length of every token"""
        n_tok, _, tok_len = self.formatter_arrays()[:3]
        return tok_len[:n_tok].astype(np.int64)

    def formatter_arrays(self):
        """(n_tokens, token pointers, token lengths, chromosome pointers, chromosome lengths, longest
        chromosome) for crp_format_rows: per-token arrays that only ever get entries appended."""
        f = self.__dict__.setdefault("_fmt", {"n": 0, "keep": [], "tok_ptr": np.zeros(16, np.uint64),
                                              "tok_len": np.zeros(16, np.uint64), "chrom_ptr": np.zeros(16, np.uint64),
                                              "chrom_len": np.zeros(16, np.uint32), "chrom_max": 0})
        n_tok = len(self.tokens)
        if n_tok > len(f["tok_ptr"]):
            for name in ("tok_ptr", "tok_len", "chrom_ptr", "chrom_len"):
                grown = np.zeros(max(n_tok, 2 * len(f[name])), f[name].dtype)
                grown[:f["n"]] = f[name][:f["n"]]
                f[name] = grown
        for k in range(f["n"], n_tok):
            view = np.frombuffer(self.tokens[k], dtype=np.uint8)
            chrom = np.frombuffer(self.chroms[k].encode("utf-8") + b"\0", dtype=np.uint8)
            f["keep"].append((view, chrom))
            f["tok_ptr"][k] = view.ctypes.data if len(view) else 0
            f["tok_len"][k] = len(view)
            f["chrom_ptr"][k] = chrom.ctypes.data
            f["chrom_len"][k] = len(chrom) - 1
            f["chrom_max"] = max(f["chrom_max"], len(chrom) - 1)
        f["n"] = n_tok
        return n_tok, f["tok_ptr"], f["tok_len"], f["chrom_ptr"], f["chrom_len"], f["chrom_max"]

    def _reserve(self, n):
        if n <= self._cap:
            return
        cap = max(n, 2 * self._cap, 1024)
        for name, dt in self._COLS:
            grown = np.empty(cap, dt)
            grown[:self.n] = self._store[name][:self.n]
            self._store[name] = grown
        self._cap = cap

    def append_token(self, key, token_bytes, seg_index, t_plus, x_plus, t_minus, x_minus):
        k = len(self.tokens)
        self.tokens.append(token_bytes)
        self.chroms.append(key[1:])
        n_p, n_m = len(t_plus), len(t_minus)
        lo, mid, hi = self.n, self.n + n_p, self.n + n_p + n_m
        self._reserve(hi)
        st = self._store
        st["tok"][lo:hi] = k
        st["seg"][lo:hi] = seg_index
        st["minus"][lo:mid] = False
        st["minus"][mid:hi] = True
        st["t"][lo:mid] = t_plus
        st["t"][mid:hi] = t_minus
        st["x"][lo:mid] = x_plus if x_plus is not None else np.nan
        st["x"][mid:hi] = x_minus if x_minus is not None else np.nan
        self.n = hi


def long_length(table, idx):
    """len(long_sequence) of candidates idx -- an index array or a slice -- (vectorised Python-slice
    arithmetic)."""
    l = table.guide_len
    t = table.t[idx].astype(np.int64)
    L = table.token_len[table.tok[idx]]
    minus = table.minus[idx]
    # '+': [t - l - 5, t + 5), '-': [t - 2, t + l + 8), clipped to the token
    lo = t - np.where(minus, 2, l + 5)
    hi = t + np.where(minus, l + 8, 5)
    np.minimum(hi, L, out=hi)
    np.maximum(lo, 0, out=lo)
    hi -= lo
    return hi


def slice_scores(table, genome, start, count, blas_threads=1):
    """on_site_score of rows [start, start+count) scored as ONE reference
    rs1_score call (CROPSR.py:461): x from the scan, the rows OpenBLAS sums in a
    non-canonical lane order re-evaluated on the GPU, then the reference's
    logistic.  Rows whose long_sequence is not 30 long get NaN (score -1 rows)."""
    idx = np.arange(start, start + count)
    x = table.x[start:start + count].copy()
    scored = long_length(table, slice(start, start + count)) == 30
    fix = {}
    if table.guide_len != 20:
        # a window truncated by the token end to exactly 30 bases is scored by the
        # reference even though guide_len != 20; its bases are those of the
        # guide_len == 20 window at a shifted position
        for i in np.nonzero(scored)[0]:
            fix[int(i)] = blas_order.CANONICAL
    for i, cls in blas_order.slice_classes(count, blas_threads).items():
        if scored[i]:
            fix[i] = cls
    if fix:
        rows = np.array(sorted(fix), dtype=np.int64)
        cand = idx[rows]
        t = table.t[cand].astype(np.int64)
        if table.guide_len != 20:
            L = table.token_len[table.tok[cand]]
            t = np.where(table.minus[cand], t, L - 5)
        strand = np.where(table.minus[cand], b"-", b"+").astype("S1")
        cls = np.array([fix[int(i)] for i in rows], dtype=np.uint8)
        x[rows] = genome.rescore(table.seg[cand], t, strand, cls)
    x[~scored] = np.nan
    with np.errstate(all="ignore"):
        return 1 / (1 + np.exp(x)), scored


def id_bytes_of(ids):
    """get_id()'s (size, 7) '<U1' array as (size, 7) ASCII bytes for the row formatter."""
    return np.ascontiguousarray(ids).view(np.uint32).astype(np.uint8)


def format_rows(table, ids, scores, scored, start, count, n_threads=0, buffer=0):
    """CSV bytes of one emitted slice through the library's multi-threaded row formatter
    (csrc/emit_csv.cpp) -- byte-identical to csv.writer().writerows() of the reference's row tuples (tests/helpers.py slice_rows)."""
    import ctypes as C
    from ._native import lib, check
    if count == 0:
        return b""
    id_bytes = ids if ids.dtype == np.uint8 else id_bytes_of(ids)
    size = len(id_bytes)
    id_index = ((start - 1 - np.arange(count, dtype=np.int64)) % size).astype(np.uint64)   # ids[start - k - 1]
    tok = np.ascontiguousarray(table.tok[start:start + count], dtype=np.uint32)
    t = np.ascontiguousarray(table.t[start:start + count], dtype=np.uint32)
    minus = np.ascontiguousarray(table.minus[start:start + count], dtype=np.uint8)
    ok = np.ascontiguousarray(scored, dtype=np.uint8)
    sc = np.ascontiguousarray(scores, dtype=np.float64)
    n_tok, tok_ptr, tok_len, chrom_ptr, chrom_len, chrom_max = table.formatter_arrays()
    # One call formats at most _FORMAT_BUDGET bytes worth of rows (the library's scratch regions are sized for the
    # worst row); the output lands in a buffer that is kept between calls -- a fresh 100+ MB buffer
    # per 1M-row slice costs more in page faults than the formatting itself.  The memoryview that is
    # returned is only valid until the next call with the same `buffer` (0 or 1).
    row_bound = 256 + 4 * table.guide_len + 2 * chrom_max
    step = max(4096, _FORMAT_BUDGET // row_bound)
    pieces = []
    for lo in range(0, count, step):
        n = min(step, count - lo)
        cap = n * row_bound
        for _ in range(2):
            if _OUT[buffer] is None or len(_OUT[buffer]) < cap:
                _OUT[buffer] = np.empty(cap, dtype=np.uint8)
            out = _OUT[buffer]
            need = C.c_uint64(0)
            rc = lib.crp_format_rows(n, id_bytes.ctypes.data, id_index[lo:].ctypes.data, tok[lo:].ctypes.data,
                                     t[lo:].ctypes.data, minus[lo:].ctypes.data, ok[lo:].ctypes.data, sc[lo:].ctypes.data,
                                     n_tok, tok_ptr.ctypes.data, tok_len.ctypes.data, chrom_ptr.ctypes.data,
                                     chrom_len.ctypes.data,
                                     int(table.guide_len), int(n_threads), out.ctypes.data, len(out), C.byref(need))
            if rc == -5:
                cap = need.value
                continue
            check(rc)
            break
        else:
            raise RuntimeError("crp_format_rows: capacity negotiation failed")
        if n == count:
            return out[:need.value].data    # a memoryview: no copy on the way to f.write()
        pieces.append(out[:need.value].tobytes())
    return b"".join(pieces)


_OUT = [None, None]
_FORMAT_BUDGET = 512 << 20       # a 1,000,000-row slice of the reference's chunk plan is one call


def write_header(path):
    with open(path, "w", newline="") as f:
        csv.writer(f).writerow(HEADER)


def emit_cumulative(path, table, genome, blas_threads=1, id_stream=None, chunk_rows=None):
    """Append the rows the reference writes after one more token has been
    scanned: ALL candidates accumulated so far (CROPSR.py:407,442), through the
    chunk plan.  Returns the number of rows written."""
    size = len(table)
    written = 0
    # the file write of slice k (0.8 s of 3.3 s on a 4 GB CSV) runs on a helper thread while slice
    # k + 1 is scored and formatted into the other output buffer; both release the GIL
    from concurrent.futures import ThreadPoolExecutor
    with open(path, "ab") as f, ThreadPoolExecutor(max_workers=1) as writer:
        ids = id_stream.next(size) if id_stream else legacy_id_bytes(size)   # get_id(size), CROPSR.py:448
        pending = [None, None]
        for k, (start, count) in enumerate(emission_slices(size, chunk_rows)):
            scores, scored = slice_scores(table, genome, start, count, blas_threads)
            if pending[k & 1] is not None:
                pending[k & 1].result()                     # that buffer's previous rows are on their way to disk
            pending[k & 1] = writer.submit(f.write, format_rows(table, ids, scores, scored, start, count, buffer=k & 1))
            written += count
        for p in pending:
            if p is not None:
                p.result()
    return written
