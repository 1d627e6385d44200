"""Primer enumeration: oracle vs the vectors the reference's own prmrdsgn2.py
functions produced (CPU), device kernel vs those vectors and vs the oracle on genome windows (GPU)."""
import hashlib
import json
import os

import numpy as np
import pytest

import primer_oracle as po

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def eng(built_lib):
    from cropsr_b200 import engine
    engine.init(0)
    return engine


def _cases():
    with open(os.path.join(HERE, "golden", "primer_vectors.json")) as f:
        return json.load(f)["cases"]


def test_oracle_matches_reference_vectors():
    """get_primers / filter_primers / Primer / pairing of prmrdsgn2.py, three parameter sets."""
    cases = _cases()
    assert len(cases) >= 25
    for c in cases:
        p, frag = c["params"], c["fragment"]
        d = po.design(frag, **p) if c["n_pairs"] < 200000 else None
        fast = po.design_fast(frag, **p)
        assert (fast["n_fwd"], fast["n_rev"], fast["n_pairs"]) == (c["n_fwd"], c["n_rev"], c["n_pairs"])
        assert (list(fast["first"]) if fast["first"] else None) == c["first"]
        if d is not None:
            assert (len(d["fwd"]), len(d["rev"]), d["n_pairs"]) == (c["n_fwd"], c["n_rev"], c["n_pairs"])
            assert (list(d["first"]) if d["first"] else None) == c["first"]
            h = hashlib.sha256()                      # sequences, GC % and Tm of every passing primer, as the reference has them
            rc = po.reverse_complement(frag)
            for seq, lst in ((frag, d["fwd"]), (rc, d["rev"])):
                for i, n in lst:
                    s = seq[i:i + n]
                    h.update(repr((s, po.gc_percentage(s), po.melting_temp(s))).encode())
            assert h.hexdigest() == c["sha256_passing"]


@pytest.mark.gpu
def test_device_primers_match_reference_vectors(eng):
    """Every golden fragment as one token; the whole token, then the fragment embedded at odd offsets."""
    from cropsr_b200 import engine, primers
    for c in _cases():
        p, frag = c["params"], c["fragment"]
        for pad_l, pad_r in ((0, 0), (37, 5), (16384 - 50, 9)):        # the last one straddles a tile edge
            tok = "T" * pad_l + frag + "A" * pad_r
            g = engine.Genome()
            g.add_token(np.frombuffer(tok.encode("latin-1"), dtype=np.uint8))
            g.commit()
            try:
                out = primers.design_windows(g, [0], [pad_l], [pad_l + len(frag)], **p)
            finally:
                g.free()
            assert out["status"][0] == 0
            assert (int(out["n_fwd"][0]), int(out["n_rev"][0]), int(out["n_pairs"][0])) == (c["n_fwd"], c["n_rev"], c["n_pairs"])
            want = c["first"] if c["first"] else [0xFFFF] * 4
            assert out["first"][0].tolist() == want


@pytest.mark.gpu
def test_device_primers_on_candidate_flanks_match_oracle(eng):
    """The +-L flank windows of real candidates (ScanResult.extras) of a multi-record genome with
    lower-case blocks and N; windows clipped at token ends that get shorter than e + l report status 1."""
    from helpers import synthetic_fasta
    from cropsr_b200 import ingest, pipeline, primers
    text = synthetic_fasta(5, [30000, 20000, 150], gc=0.48, lower_frac=0.3, n_frac=0.003)
    tokens = ingest.fasta_text_to_tokens(text)
    genome, result, _ = pipeline.scan_tokens(tokens, 20)
    try:
        rng = np.random.default_rng(3)
        for seg, (key, tok) in enumerate(tokens.items()):
            for strand in "+-":
                ex = result.extras(seg, strand, flank=200)
                n = len(ex["cut"])
                if n == 0:
                    continue
                pick = np.unique(np.concatenate((np.arange(min(n, 6)), np.arange(max(n - 6, 0), n), rng.integers(0, n, 25))))
                lo, hi = ex["flank_lo"][pick], ex["flank_hi"][pick]
                out = primers.design_windows(genome, np.full(len(pick), seg), lo, hi)
                for k in range(len(pick)):
                    frag = tok[int(lo[k]):int(hi[k])]
                    if len(frag) < 130:
                        assert out["status"][k] == 1
                        continue
                    want = po.design_fast(frag)
                    assert out["status"][k] == 0
                    assert (int(out["n_fwd"][k]), int(out["n_rev"][k]), int(out["n_pairs"][k])) == \
                        (want["n_fwd"], want["n_rev"], want["n_pairs"])
                    assert out["first"][k].tolist() == (list(want["first"]) if want["first"] else [0xFFFF] * 4)
    finally:
        result.free()
        genome.free()


@pytest.mark.gpu
def test_device_primers_reject_bad_arguments(eng):
    from cropsr_b200 import engine, primers, _native
    g = engine.Genome()
    g.add_token(np.frombuffer(b"ACGT" * 100, dtype=np.uint8))
    g.commit()
    try:
        with pytest.raises(_native.CropsrError):
            primers.design_windows(g, [0], [0], [401])               # beyond the token
        with pytest.raises(_native.CropsrError):
            primers.design_windows(g, [1], [0], [400])               # no such segment
        with pytest.raises(_native.CropsrError):
            primers.design_windows(g, [0], [0], [400], s=5)          # N < 13 would need the other Tm formula
        with pytest.raises(TypeError):
            primers.design_windows(g, [0], [0], [400], tm=3)
        out = primers.design_windows(g, [0], [0], [100])             # shorter than e + l
        assert out["status"][0] == 1 and out["n_pairs"][0] == 0
    finally:
        g.free()
