"""End-to-end --cas9 run: ingest -> GPU scan+score -> reference-order emission.

Mirror of /root/reference/CROPSR.py:333-486 main().  Differences that are
deliberate and documented in DESIGN.md: the whole genome is scanned on the GPU
in one launch sequence before the per-token emission loop starts, and the
reference's 5-second sleep per token (CROPSR.py:478) is not performed.
"""
import time

import numpy as np

from . import emit, engine, ingest


def scan_tokens(tokens, guide_len=20, flags=0):
    """Pack every token of the ingest dict into HBM and scan it.
    Returns (genome, result, token_bytes)."""
    genome = engine.Genome()
    token_bytes = []
    for value in tokens.values():
        b = value.encode("ascii") if isinstance(value, str) else bytes(value)
        token_bytes.append(b)
        genome.add_token(b)
    genome.commit()
    result = genome.scan(guide_len, flags)
    return genome, result, token_bytes


def run_cas9(fasta, gff, output="data.csv", guide_len=20, verbose=False, blas_threads=1,
             time_path="time.txt", out=print):
    begin = time.time()
    timing = open(time_path, "w")                       # CROPSR.py:371
    tokens = ingest.import_fasta_file(fasta, verbose)   # :374
    ingest.import_gff_file(gff, verbose)                # :375 (parsed, never used)
    if verbose:
        out("\n            Initiating PAM site detection.\n            \n"
            "            Please wait, this may take a while...\n            ")
    emit.write_header(output)                           # :402-405

    genome, result, token_bytes = scan_tokens(tokens, guide_len)
    table = emit.CandidateTable(guide_len)
    rows_written = 0
    for seg, (key, value) in enumerate(tokens.items()):
        out("Searching on Chromosome: ", key[:25])       # :410-411
        out("With start of sequence: ", value[:25])
        plus = result.fetch_segment(seg, "+", want=("pos", "x"))
        minus = result.fetch_segment(seg, "-", want=("pos", "x"))
        table.append_token(key, token_bytes[seg], seg, plus["pos"], plus["x"], minus["pos"], minus["x"])
        if verbose:
            n = len(plus["pos"]) + len(minus["pos"])
            out(f"\n                {n:n} Cas9 PAM sites were found on {key[1:]}\n                ")
        rows_written += emit.emit_cumulative(output, table, genome, blas_threads)
        timing.write("Total runtime of the program is " + str(time.time() - begin))   # :476-477
    timing.close()
    stats = {"tokens": len(tokens), "candidates": len(table), "rows": rows_written,
             "scan_ms": result.scan_ms(), **genome.timing()}
    result.free()
    genome.free()
    if verbose:
        out(f"The output file has been generated at {output}")
    return stats
