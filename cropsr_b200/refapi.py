"""The importable callables of the reference's CROPSR.py that sit on the Cas9 path, under their
own names and signatures -- what `import CROPSR` must keep offering.

    import_fasta_file  CROPSR.py:54-74      import_gff_file   :77-95
    find_PAM_site      :98-104              get_reverse_complement / get_gRNA_sequence  :116-129
    apply_cutsite      :155-158             rs1_score         :285-313       get_id  :316-318

rs1_score and find_PAM_site compute on the GPU through libcropsr_b200 (no CPU fallback: they
raise if the library cannot drive a device); the string transforms, ids and file readers are the
host logic they are in the reference.  Unlike the reference module, importing this one does not
parse sys.argv.
"""
import numpy as np

from . import blas_order
from .emit import alphanum, apply_cutsite, get_id                      # noqa: F401
from .ingest import import_fasta_file, import_gff_file                 # noqa: F401

PAM_PLUS, PAM_MINUS = "(?=.GG)", "(?=CC.)"            # CROPSR.py:415, :426


def get_reverse_complement(input_sequence):
    """CROPSR.py:116-121, the literal replace chain."""
    return (input_sequence.replace("A", "U").replace("C", "Z").replace("G", "C").replace("Z", "G")
            .replace("T", "A").replace("U", "T")[::-1])


def get_gRNA_sequence(input_sequence):
    """CROPSR.py:124-129, the literal replace chain."""
    return input_sequence.replace("A", "U").replace("C", "Z").replace("G", "C").replace("Z", "G").replace("T", "A")[::-1]


def find_PAM_site(target, input_sequence):
    """[(t, t), ...] for every overlapping match of the Cas9 PAM look-ahead `target` in
    input_sequence, ascending (CROPSR.py:98-104 with the two patterns main() passes, :415/:426).
    The scan runs on the GPU with the window bounds switched off as far as the kernel allows
    (guide length 1: '+' from t = 6, '-' from t = 2); the handful of positions before that are
    compared here.  Other regular expressions are not part of the accelerated path."""
    from . import engine, _native as N
    if target not in (PAM_PLUS, PAM_MINUS):
        raise NotImplementedError(f"find_PAM_site: only {PAM_PLUS!r} and {PAM_MINUS!r} (the patterns CROPSR.py scans "
                                  f"for) run on the device; got {target!r}")
    tok = input_sequence.encode("ascii") if isinstance(input_sequence, str) else bytes(input_sequence)
    minus = target == PAM_MINUS
    L = len(tok)
    head = []
    for t in range(min(2 if minus else 6, max(L - 2, 0))):
        if (tok[t:t + 2] == b"CC") if minus else (tok[t + 1:t + 3] == b"GG"):
            head.append(t)
    body = np.empty(0, np.uint32)
    if L:
        genome = engine.Genome()
        try:
            genome.add_token(tok)
            genome.commit()
            result = genome.scan(1, N.CRP_SCAN_NO_SCORE)
            try:
                body = result.fetch_segment(0, "-" if minus else "+", want=("pos",))["pos"]
            finally:
                result.free()
        finally:
            genome.free()
    return [(int(t), int(t)) for t in head] + [(int(t), int(t)) for t in body]


def rs1_score(sequences, blas_threads=1):
    """CROPSR.py:285-313 on the device: Rule Set 1 of every row of the (n, 30) byte matrix, each
    row summed in the lane order OpenBLAS gives it inside an n-row np.matmul (blas_order), then
    1 / (1 + np.exp(.)) with numpy's digits.  blas_threads: the OpenBLAS thread count to emulate."""
    from . import engine
    seqs = np.asarray(sequences)
    if seqs.ndim != 2 or seqs.shape[1] != 30:
        raise ValueError("rs1_score expects an (n, 30) matrix of ASCII codes")
    if seqs.dtype != np.uint8:                        # placeholder rows make the reference's matrix float64
        ok = np.isin(seqs, (65, 84, 67, 71))
        seqs = np.where(ok, seqs, 0).astype(np.uint8)
    n = len(seqs)
    cls = np.zeros(n, dtype=np.uint8)
    for i, c in blas_order.slice_classes(n, blas_threads).items():
        cls[i] = c
    return engine.rs1_score(np.ascontiguousarray(seqs), cls)
