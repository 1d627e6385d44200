"""The oracle is only trustworthy if it reproduces the unmodified reference:
byte-for-byte against the golden CSVs made by tests/golden/make_golden.py and
against the SURVEY section-4 digests of the reference's shipped output.csv."""
import hashlib
import os

import numpy as np
import pytest

import cropsr_oracle as oracle
import rs1_table
from helpers import fixture_text, golden_csv, normalised_digest

# digests of /root/reference/sample_data/output.csv
SHIPPED_DIGEST_NO_SCORE = "16c177dd33f76fe7ff47c5953efa3c3ba4ce1c8c5f4c43d040bdc617edf80c9d"
SHIPPED_DIGEST_SCORE_12G = "4cc3ee77c556c5b3e1df05912306631193fefeeeacf55762eb9e9dc08af5ae3b"


def test_weight_table_digests():
    rs1_table.check_digests()


def test_manifest_matches_files(manifest):
    for name, case in manifest["cases"].items():
        assert hashlib.sha256(golden_csv(name).encode()).hexdigest() == case["csv_sha256"], name


@pytest.mark.parametrize("name", [
    "sample", "sample_t6", "multi3", "multi3_l18", "multi3_l23", "clean3", "clean3_trailing_nl",
    "edge_clean", "edge_fmt", "single_candidate", "ws_header", "dup_keys", "empty_records",
    "mid50k", "mid50k_t5", "mid50k_c1000", "mid50k_c1557", "multi3_c20", "sample_c5000_t4"])
def test_oracle_reproduces_reference_csv(name, manifest):
    """The *_c<chunk> cases are the reference run with its 1000000-row chunk literal shrunk
    (tests/golden/make_golden.py): the chunk plan of CROPSR.py:451-472 on small inputs."""
    case = manifest["cases"][name]
    np.random.seed(case["seed"])
    got = oracle.run_to_string(fixture_text(case["fasta"]), case["guide_len"], "model", case["blas_threads"],
                               chunk=case.get("chunk"))
    assert got == golden_csv(name)


def test_sample_matches_shipped_output_digests():
    text = golden_csv("sample")
    assert normalised_digest(text) == SHIPPED_DIGEST_NO_SCORE
    assert normalised_digest(text, "%.12g") == SHIPPED_DIGEST_SCORE_12G


def test_thread_model_matters():
    # the multi-thread goldens differ from the 1-thread ones only in a few score strings
    assert golden_csv("mid50k") != golden_csv("mid50k_t5")
    assert normalised_digest(golden_csv("mid50k")) == normalised_digest(golden_csv("mid50k_t5"))


def test_model_order_equals_this_machines_blas():
    """score_model (explicit lane order) vs score_blas (np.matmul) -- only
    meaningful where numpy links the OpenBLAS build the model was derived on."""
    import os
    if os.environ.get("OPENBLAS_NUM_THREADS", "") != "1":
        pytest.skip("needs OPENBLAS_NUM_THREADS=1 set before numpy is imported")
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 4, 5, 6, 7, 50, 1003):
        seqs = rng.choice(np.frombuffer(b"ATCGN", dtype=np.uint8), size=(n, 30))
        a, b = oracle.score_blas(seqs), oracle.score_model(seqs, 1)
        if not np.array_equal(a, b):
            pytest.skip("this machine's BLAS sums in a different order than the build container's")


def test_emission_slices_closed_form():
    for n in list(range(0, 40)) + [999999, 1000000, 1000001, 1124799, 2000000, 2000001, 3000000, 3500007]:
        assert oracle.emission_slices(n) == oracle.emission_slices_closed_form(n)
    # 1,124,799 rows -> rows 0..999,999 then 124,799..249,597
    assert oracle.emission_slices(1124799) == [(0, 1000000), (124799, 124799)]
    assert oracle.emission_slices(1000000) == []


def _np_exp_oracle():
    import ctypes
    lib_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "libnp_exp.so")
    if not os.path.exists(lib_path):
        import subprocess
        subprocess.run(["make", "-C", os.path.dirname(lib_path)], check=True)
    return ctypes.CDLL(lib_path)


def oracle_np_exp(x):
    import ctypes
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    _np_exp_oracle().np_exp_f64_array(x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p),
                                       ctypes.c_long(x.size))
    return y


def test_np_exp_oracle_equals_the_golden_vectors():
    """oracle/np_exp.c (numpy's SVML exp restated) against vectors produced by the reference
    environment's own np.exp (tests/golden/make_np_exp_vectors.py): bit for bit."""
    v = np.load(os.path.join(os.path.dirname(__file__), "golden", "np_exp_vectors.npz"))
    got = oracle_np_exp(v["x"])
    assert np.array_equal(got, v["exp"])
    with np.errstate(all="ignore"):
        assert np.array_equal(1.0 / (1.0 + got), v["score"])


def test_np_exp_oracle_equals_numpy_on_avx512_hosts():
    """...and against np.exp itself where numpy takes the AVX-512 SVML path (the golden host did)."""
    feats = getattr(getattr(np, "_core", None), "_multiarray_umath", None)
    feats = getattr(feats, "__cpu_features__", {}) if feats is not None else {}
    if not feats.get("AVX512_SKX"):
        pytest.skip("numpy does not dispatch exp to the AVX-512 SVML routine on this host")
    rng = np.random.default_rng(8)
    x = np.concatenate([rng.uniform(-18, 9, 400000), rng.uniform(-700, 700, 100000), rng.normal(0, 1, 100001)])
    assert np.array_equal(oracle_np_exp(x), np.exp(x))
    assert np.array_equal(oracle_np_exp(x[:13]), np.exp(x[:13]))        # numpy's tail loop is the same routine


def test_every_scored_window_carries_the_pam_as_cc():
    """csrc/gen_rs1_inc.py folds the PAM into the Rule-Set-1 lane tables: every 30-mer the
    reference scores has upper-case C at 0-based positions 2 and 3 on both strands
    (CROPSR.py:415-433).  Pinned here on the oracle's candidates of every fixture."""
    seen = 0
    for name in ("sample_genome.fa", "multi3.fa", "edge_fmt.fa", "edge_clean.fa", "mid50k.fa", "clean3.fa"):
        from cropsr_b200 import ingest
        for key, tok in ingest.fasta_text_to_tokens(fixture_text(name)).items():
            for cand in oracle.candidates_for_token(key, tok, 20):
                long_ = cand[4]
                if len(long_) == 30:
                    assert long_[2:4] == "CC", (name, cand)
                    seen += 1
    assert seen > 20000


@pytest.mark.slow
@pytest.mark.parametrize("name", ["big1", "big2"])
def test_oracle_reproduces_reference_over_one_million_candidates(name):
    """The literal port against the digests of the unmodified reference's >1e6-candidate runs
    (tests/golden/make_big_golden.py); minutes of Python, so CROPSR_SLOW=1 only."""
    import importlib.util, json
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_big_golden", os.path.join(here, "make_big_golden.py"))
    mbg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mbg)
    with open(os.path.join(here, "cases", "big_manifest.json")) as f:
        case = json.load(f)["cases"][name]
    np.random.seed(case["seed"])
    got = oracle.run_to_string(mbg.big_fasta(name), 20, "model", case["blas_threads"])
    assert hashlib.sha256(got.encode()).hexdigest() == case["csv_sha256"]


def _check_vector_oracle(text, guide_len=20):
    import vector_oracle as vo
    tokens = oracle.import_fasta_text(text)
    for key, tok in tokens.items():
        arr = np.frombuffer(tok.encode("ascii"), dtype=np.uint8)
        tab = vo.token_table(arr, guide_len)
        plus, minus = oracle.pam_hits(tok, guide_len)
        assert tab["+"][0].tolist() == plus and tab["-"][0].tolist() == minus
        if guide_len != 20:
            continue
        cands = oracle.candidates_for_token(key, tok, guide_len)
        packed = np.concatenate((tab["+"][1], tab["-"][1]))
        x = np.concatenate((tab["+"][2], tab["-"][2]))
        full = np.array([len(c[4]) == 30 for c in cands], dtype=bool)
        assert np.array_equal((packed & np.uint64(1 << 31)) != 0, ~full)
        if not full.any():
            continue
        seqs = np.array([oracle.scored_bytes(c[4]) for c, f in zip(cands, full) if f])
        assert np.array_equal(x[full], oracle.preactivation_model(seqs, classes=np.zeros(len(seqs), dtype=np.int8)))
        code = np.zeros(seqs.shape, dtype=np.uint64)
        scoring = np.zeros(seqs.shape, dtype=bool)
        for c, b in enumerate(b"ATCG"):
            code[seqs == b] = c
            scoring |= seqs == b
        sh = np.arange(30, dtype=np.uint64)
        lo = ((code & np.uint64(1)) << sh).sum(axis=1, dtype=np.uint64)
        hi = ((code >> np.uint64(1)) << sh).sum(axis=1, dtype=np.uint64)
        pk = packed[full]
        assert np.array_equal(pk & np.uint64(0x3FFFFFFF), lo) and np.array_equal((pk >> np.uint64(32)) & np.uint64(0x3FFFFFFF), hi)
        assert np.array_equal((pk & np.uint64(1 << 62)) != 0, ~scoring.all(axis=1))


@pytest.mark.parametrize("fasta", ["multi3.fa", "clean3.fa", "edge_clean.fa", "edge_fmt.fa", "ws_header.fa", "dup_keys.fa",
                                   "empty_records.fa", "single_candidate.fa", "mid50k.fa", "sample_genome.fa"])
@pytest.mark.parametrize("guide_len", [20, 18, 23])
def test_vector_oracle_equals_literal_port(fasta, guide_len):
    """oracle/vector_oracle.py (numpy, from the spec) against the literal port that the reference's CSVs pin:
    positions, truncation, scored codes, fp64 x -- so that its digests can stand in at 135 Mbp - 10 Gbp."""
    _check_vector_oracle(fixture_text(fasta), guide_len)


@pytest.mark.parametrize("seed", range(6))
def test_vector_oracle_on_random_fastas(seed):
    rng = np.random.default_rng(2000 + seed)
    alphabet = np.frombuffer(b"ACGTacgtNRYUZuz'", dtype=np.uint8)
    p = np.array([20, 20, 20, 20, 3, 3, 3, 3, 1, .3, .3, .2, .2, .1, .1, .1])
    recs = [(f"r{k}", rng.choice(alphabet, size=int(rng.integers(0, 6000)), p=p / p.sum()).tobytes().decode()) for k in range(4)]
    if seed % 2:
        text = "".join(f">{h}\n" + "\n".join(s[i:i + 70] for i in range(0, len(s), 70)) + "\n" for h, s in recs)
    else:
        text = "\n".join(f">{h}\n{s}" for h, s in recs if s)
    _check_vector_oracle(text)
