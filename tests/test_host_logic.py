"""Host-side logic of the product package, checked against the oracle on the CPU
(no compute calls into the CUDA library)."""
import glob
import os
import re

import numpy as np
import pytest

import cropsr_oracle as oracle
from helpers import GOLDEN, ROOT, fixture_text, ids_to_strings, slice_rows


def test_library_loads_and_exports_header_symbols(built_lib):
    header = open(os.path.join(ROOT, "include", "cropsr_b200.h")).read()
    declared = set(re.findall(r"\b(crp_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(built_lib.SIGNATURES), declared ^ set(built_lib.SIGNATURES)
    for name in declared:
        getattr(built_lib.lib, name)
    # the dynamic symbol table of the .so itself: every crp_* it exports is declared, and nothing else is C-named
    import subprocess
    nm = subprocess.run(["nm", "-D", "--defined-only", built_lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in nm.splitlines() if ln.split()[-1].startswith("crp_")}
    assert exported == declared, exported ^ declared
    assert built_lib.lib.crp_abi_version() == built_lib.ABI_VERSION
    assert built_lib.lib.crp_last_error() is not None


def test_compute_fails_loudly_without_gpu(built_lib):
    import ctypes as C
    n = C.c_int(-1)
    rc = built_lib.lib.crp_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    from cropsr_b200 import engine
    with pytest.raises(built_lib.CropsrError):
        engine.init(0)
    g = C.c_void_p()
    assert built_lib.lib.crp_genome_new(C.byref(g)) != 0       # no silent CPU path
    assert b"crp_init" in built_lib.lib.crp_last_error()


def test_ingest_matches_oracle(built_lib):
    from cropsr_b200 import ingest
    for path in glob.glob(os.path.join(GOLDEN, "fixtures", "*.fa")):
        text = fixture_text(os.path.basename(path))
        assert ingest.formatted(text) == oracle.formatted(text)
        assert list(ingest.fasta_text_to_tokens(text).items()) == list(oracle.import_fasta_text(text).items())
    for text in ("", ">a", ">a\n", ">a\nACGT", ">a\nACGT\n>b\nGG", "no header at all\nACGT\n", ">x y\nAC\nGT\n"):
        assert list(ingest.fasta_text_to_tokens(text).items()) == list(oracle.import_fasta_text(text).items())


def test_drop_in_module_names(built_lib):
    import cropsr_functions
    assert cropsr_functions.formatted(">a\nAC\nGT\n") == "[('a', 'ACGT')]"
    assert cropsr_functions.generate_dictionary("k1 v1 k2") == {"k1": "v1", "k2": ""}


def test_guide_strings_match_reference_transforms(built_lib):
    from cropsr_b200 import emit
    for fasta in ("edge_clean.fa", "edge_fmt.fa", "multi3.fa", "clean3.fa"):
        for key, tok in oracle.import_fasta_text(fixture_text(fasta)).items():
            for l in (18, 20, 21, 22, 23, 30):
                for c in oracle.candidates_for_token(key, tok, l):
                    minus = c[6] == "-"
                    t = c[1] - 3 if minus else c[1]
                    assert emit.guide_strings(tok.encode(), t, minus, l) == (c[3], c[4])


def test_emission_slices(built_lib):
    from cropsr_b200 import emit
    for n in list(range(0, 40)) + [999999, 1000000, 1000001, 1124799, 2000000, 2000001, 3000000, 3500007]:
        assert emit.emission_slices(n) == oracle.emission_slices(n)


def test_blas_row_classes_match_oracle_model(built_lib):
    from cropsr_b200 import blas_order
    for threads in (1, 2, 8, 64):
        for n in list(range(1, 70)) + [993, 994, 1001, 3839, 3840, 3842, 5003, 12345, 99998, 250003, 1000000]:
            for d in (120, 464):
                want = {i: int(c) for i, c in enumerate(oracle.row_classes(n, d, threads)) if c}
                assert blas_order.noncanonical_rows(n, d, threads) == want, (threads, n, d)


def test_ids_reproduce_reference_rng(built_lib):
    from cropsr_b200 import emit
    np.random.seed(3)
    a = ids_to_strings(emit.get_id(1000))
    np.random.seed(3)
    b = oracle.make_ids(1000)
    assert a == b and len(a[0]) == 7


def test_float_repr_equals_numpy_str():
    rng = np.random.default_rng(0)
    v = np.concatenate([1 / (1 + np.exp(rng.uniform(-20, 10, 200000))), 10.0 ** rng.uniform(-300, 300, 20000)])
    assert [repr(x) for x in v.tolist()] == [str(x) for x in v]


def test_shard_plan_and_offsets(built_lib):
    from cropsr_b200 import shard
    rng = np.random.default_rng(1)
    for _ in range(50):
        lens = [int(x) for x in rng.integers(0, 200000, size=int(rng.integers(1, 8)))]
        for ws in (1, 2, 3, 4, 8):
            plans = shard.plan(lens, ws)
            # every position owned exactly once, in order, boundaries on the tile granule
            seen = []
            for segs in plans:
                for k, a, b in segs:
                    assert a % shard.TILE == 0 and 0 <= a <= b <= lens[k]
                    seen.append((k, a, b))
            assert seen == sorted(seen)
            for k, n in enumerate(lens):
                pieces = [(a, b) for kk, a, b in seen if kk == k and b > a]
                assert sum(b - a for a, b in pieces) == n
                for (a0, b0), (a1, b1) in zip(pieces, pieces[1:]):
                    assert b0 == a1
            # offsets: simulate counts = number of "hits" at fake positions
            counts = [([(b - a) // 7 for _, a, b in segs], [(b - a) // 11 for _, a, b in segs]) for segs in plans]
            offs, total = shard.global_offsets(plans, counts)
            assert total == sum(sum(c[0]) + sum(c[1]) for c in counts)
            rows = []
            for r, segs in enumerate(plans):
                for s, (k, a, b) in enumerate(segs):
                    rows.append((offs[r][s][0], counts[r][0][s], k, 0, a))
                    rows.append((offs[r][s][1], counts[r][1][s], k, 1, a))
            rows = [x for x in rows if x[1]]
            rows.sort()
            pos = 0
            for off, c, *_ in rows:
                assert off == pos
                pos += c
            assert [(k, st, a) for _, _, k, st, a in rows] == sorted((k, st, a) for _, _, k, st, a in rows)


def test_gff_intervals_per_token(built_lib):
    """GFF rows -> sorted inclusive token intervals, both ingest paths (annotate.py)."""
    import os
    from cropsr_b200 import annotate, ingest
    gff = os.path.join(os.path.dirname(__file__), "golden", "fixtures", "sample_genome.gff")
    frame = ingest.import_gff_file(gff)
    chrom = frame.loc[frame["feature"] == "gene", "chromosome"].iloc[0]     # the #! header lines are junk rows
    for formatted, key in ((True, f"[('{chrom}',"), (True, f"('{chrom}',"), (False, f">{chrom}")):     # first / later record
        iv, = annotate.intervals_for_tokens(frame, [key], formatted, features=("gene",))
        genes = frame[(frame["feature"] == "gene") & (frame["chromosome"] == chrom)]
        assert len(iv["start"]) == len(genes) > 0
        assert np.all(np.diff(iv["start"].astype(np.int64)) >= 0)
        shift = 0 if formatted else -1
        assert iv["start"].min() == int(genes["start"].min()) + shift
        assert set(frame.loc[iv["row"], "feature"]) == {"gene"}
    none, = annotate.intervals_for_tokens(frame, [">not_there"], False)
    assert len(none["start"]) == 0


def _table_from_oracle(text, guide_len):
    import cropsr_oracle as oracle
    from cropsr_b200 import emit, ingest
    tokens = ingest.fasta_text_to_tokens(text)
    table = emit.CandidateTable(guide_len)
    for seg, (key, tok) in enumerate(tokens.items()):
        plus, minus = oracle.pam_hits(tok, guide_len)
        table.append_token(key, tok.encode(), seg, np.array(plus, np.uint32), None, np.array(minus, np.uint32), None)
    return table


@pytest.mark.parametrize("fasta,guide_len", [("edge_fmt.fa", 20), ("edge_clean.fa", 20), ("multi3.fa", 18),
                                             ("mid50k.fa", 20), ("ws_header.fa", 20)])
def test_c_row_formatter_equals_python_csv_writer(built_lib, fasta, guide_len, monkeypatch):
    """csrc/emit_csv.cpp vs the literal Python row tuples + csv.writer: quoting of decoration
    bytes, truncated windows / 11-field error rows, repr() of the scores, id reverse indexing."""
    import csv, io
    from cropsr_b200 import emit
    from helpers import fixture_text
    table = _table_from_oracle(fixture_text(fasta), guide_len)
    n = len(table)
    assert n > 0
    rng = np.random.default_rng(3)
    np.random.seed(4)
    ids = emit.get_id(n)
    scored = emit.long_length(table, np.arange(n)) == 30
    scores = 1 / (1 + np.exp(rng.uniform(-18, 9, n)))
    scores[::7] = 10.0 ** rng.uniform(-300, 300, len(scores[::7]))         # exercise the exponent layouts too
    scores[~scored] = np.nan
    for start, count in ((0, n), (n // 3, n - n // 3), (n - 1, 1)):
        want = io.StringIO(newline="")
        csv.writer(want).writerows(slice_rows(table, ids_to_strings(ids), scores[start:start + count],
                                                   scored[start:start + count], start, count))
        for threads, budget in ((1, 512 << 20), (3, 512 << 20), (2, 1 << 20)):      # the last one: several calls per slice
            monkeypatch.setattr(emit, "_FORMAT_BUDGET", budget)
            got = emit.format_rows(table, ids, scores[start:start + count], scored[start:start + count], start, count,
                                   n_threads=threads)
            assert bytes(got) == want.getvalue().encode()


def test_c_float_repr_matches_python(built_lib):
    """repr(float) of the formatter on many values: shortest round-trip digits, fixed vs exponent."""
    import ctypes as C
    from cropsr_b200 import emit
    rng = np.random.default_rng(5)
    vals = np.concatenate([1 / (1 + np.exp(rng.uniform(-20, 10, 60000))), 10.0 ** rng.uniform(-320, 308, 30000),
                           2.0 ** rng.integers(-1074, 1023, 4000).astype(np.float64), np.ldexp(rng.random(4000), -1070),
                           np.array([0.5, 1.0, 1e16, 1e15, 9007199254740992.0, 0.0001, 0.00001, 5e-324, 1.7976931348623157e308,
                                     123456789012345.0, 1e22, 1e23, 0.1, 0.2 + 0.1])])
    table = emit.CandidateTable(20)
    tok = b"A" * 40 + b"GG" + b"A" * 40
    table.append_token(">c", tok, 0, np.full(len(vals), 39, np.uint32), None, np.empty(0, np.uint32), None)
    ids = np.full((len(vals), 7), "A", dtype="<U1")
    got = bytes(emit.format_rows(table, ids, vals, np.ones(len(vals), bool), 0, len(vals))).decode().split("\r\n")[:-1]
    assert [row.split(",")[9] for row in got] == [repr(float(v)) for v in vals]


def test_plain_fasta_layout_matches_text_ingest():
    """ingest.plain_fasta_layout (device ingest) derives the same keys and tokens as the literal
    text pipeline on every plain fixture and refuses everything the text quirks would change."""
    import glob, os
    from cropsr_b200 import ingest
    from helpers import synthetic_fasta
    fixtures = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "fixtures", "*.fa")))
    texts = {os.path.basename(f): open(f, "rb").read() for f in fixtures}
    texts["synthetic"] = synthetic_fasta(5, [1000, 77, 80, 161], gc=0.5, lower_frac=0.2).encode()
    plain = 0
    for name, data in texts.items():
        lay = ingest.plain_fasta_layout(data)
        toks = ingest.fasta_text_to_tokens(data.decode())
        if lay is None:
            continue
        plain += 1
        assert [rec[0] for rec in lay] == list(toks.keys()), name
        for (key, off, nbytes, width, last), value in zip(lay, toks.values()):
            body = data[off:off + nbytes]
            assert b"\n" not in body[:width] and (len(body) <= width or body[width:width + 1] == b"\n")
            assert value.encode() == b"'" + body.replace(b"\n", b"") + (b"')]" if last else b"'),"), name
    assert plain >= 5
    for name in ("clean3.fa", "dup_keys.fa", "ws_header.fa", "empty_records.fa", "edge_clean.fa"):
        assert ingest.plain_fasta_layout(texts[name]) is None, name
    for bad in (b"", b"ACGT\n", b">a\nAC\r\nGT\n", b">a b\nACGT\nAC\n", b">a\nACGT\n>a\nAC\nA\n", b">a\n>b\nAC\nA\n",
                b">a'\nACGT\nA\n", b">a\nAC>GT\nAA\n"):
        assert ingest.plain_fasta_layout(bad) is None, bad


def test_cropsr_module_keeps_the_reference_callables(built_lib):
    """`import CROPSR` offers the callables of the reference's module that sit on the Cas9 path
    (SURVEY 8b), and the host-side ones behave like the reference's (literal replace chains)."""
    import CROPSR
    for name in ("import_fasta_file", "import_gff_file", "find_PAM_site", "get_reverse_complement",
                 "get_gRNA_sequence", "apply_cutsite", "rs1_score", "get_id", "main"):
        assert callable(getattr(CROPSR, name)), name
    rng = np.random.default_rng(12)
    for _ in range(200):
        s = "".join(rng.choice(list("ACGTacgtNUZ'),"), size=int(rng.integers(0, 40))))
        assert CROPSR.get_gRNA_sequence(s) == oracle.grna(s)
        assert CROPSR.get_reverse_complement(s) == oracle.reverse_complement(s)
    assert CROPSR.apply_cutsite(5, 25, "cas9") == 22
    np.random.seed(3)
    a = CROPSR.get_id(5)
    np.random.seed(3)
    assert np.array_equal(a, np.random.choice(CROPSR.alphanum, [5, 7]))
    with pytest.raises(NotImplementedError):
        CROPSR.find_PAM_site("(?=TTT)", "ACGT")


def test_phytozome_annotation_info_join(tmp_path):
    """-p annotation_info.txt (echoed and ignored by the reference, CROPSR.py:32,364): rows keyed by
    pacId / locus / transcript / peptide, looked up from GFF3 attributes with version suffixes peeled."""
    from cropsr_b200 import annotate
    p = tmp_path / "Sbicolor_454_v3.1.1.annotation_info.txt"
    p.write_text("#pacId\tlocusName\ttranscriptName\tpeptideName\tPfam\tPanther\tKOG\tec\tKO\tGO\tBest-hit-arabi-name\tarabi-symbol\tarabi-defline\n"
                 "37916712\tSobic.001G000100\tSobic.001G000100.1\tSobic.001G000100.1.p\tPF00069\t\t\t2.7.11.1\t\tGO:0004672\tAT1G01540.2\t\tProtein kinase\n"
                 "37916713\tSobic.001G000200\tSobic.001G000200.2\tSobic.001G000200.2.p\t\t\t\t\t\t\t\t\t\n")
    t = annotate.read_annotation_info(str(p))
    want = "Pfam=PF00069;ec=2.7.11.1;GO=GO:0004672;Best-hit-arabi-name=AT1G01540.2;arabi-defline=Protein kinase"
    assert t["Sobic.001G000100"] == want and t["37916712"] == want and t["Sobic.001G000100.1.p"] == want
    assert t["Sobic.001G000200"] == ""
    look = annotate.lookup_annotation_info
    assert look(t, "ID=Sobic.001G000100.v3.1;Name=Sobic.001G000100") == want
    assert look(t, "ID=Sobic.001G000100.1.v3.1.CDS.2;Parent=Sobic.001G000100.1.v3.1;pacid=37916712") == want
    assert look(t, "ID=cds-1;Parent=Sobic.001G000100.1.v3.1") == want
    assert look(t, "ID=unknown.7;Name=other") == "" and look(t, float("nan")) == "" and look({}, "ID=x") == ""


def test_legacy_ids_equal_numpy_choice_and_leave_the_same_generator_state(built_lib):
    """crp_legacy_ids (the library's MT19937 loop) against get_id() = np.random.choice(alphanum, [n, 7])
    (CROPSR.py:316-318): same characters, same generator state afterwards (so every later draw is the
    same too), from arbitrary positions in the stream, across state refills, for empty requests."""
    from cropsr_b200 import emit
    for seed, skip, n in ((1, 0, 0), (2, 0, 1), (3, 5, 89), (4, 623, 90), (5, 624, 1000), (6, 17, 20011), (7, 1, 250000)):
        np.random.seed(seed)
        np.random.randint(0, 2 ** 32, size=skip, dtype=np.uint32)         # move inside the 624-word block
        want = emit.id_bytes_of(emit.get_id(n)).reshape(-1, 7)
        state_want = np.random.get_state()
        after_want = np.random.choice(emit.alphanum, [5, 7])
        np.random.seed(seed)
        np.random.randint(0, 2 ** 32, size=skip, dtype=np.uint32)
        got = emit.legacy_id_bytes(n)
        state_got = np.random.get_state()
        after_got = np.random.choice(emit.alphanum, [5, 7])
        assert got.shape == (n, 7) and np.array_equal(got, want)
        assert state_got[0] == state_want[0] and np.array_equal(state_got[1], state_want[1]) and state_got[2:] == state_want[2:]
        assert np.array_equal(after_got, after_want)


def test_id_stream_draws_ahead_what_get_id_would_draw(built_lib):
    """emit.IdStream: ids of a known sequence of emission sizes generated on a helper thread -- same
    characters as get_id(size) per emission, same generator state after close(), also when the
    stream is closed early."""
    from cropsr_b200 import emit
    sizes = [3, 10, 10, 0, 2500, 40001]
    np.random.seed(77)
    want = [emit.id_bytes_of(emit.get_id(n)).reshape(-1, 7) for n in sizes]
    after_want = np.random.randint(0, 1000, 4)
    np.random.seed(77)
    st = emit.IdStream(sizes)
    got = [st.next(n) for n in sizes]
    st.close()
    after_got = np.random.randint(0, 1000, 4)
    assert all(np.array_equal(a, b) for a, b in zip(got, want)) and np.array_equal(after_got, after_want)
    np.random.seed(77)
    st = emit.IdStream(sizes)
    assert np.array_equal(st.next(3), want[0])
    with pytest.raises(AssertionError):
        st.next(11)
    st.close()                                   # drains: the state is the one after all six draws
    assert np.array_equal(np.random.randint(0, 1000, 4), after_want)


def test_shard_plan_tiles_the_genome_and_balances_tiles(built_lib):
    """shard.plan: every position of every token in exactly one segment, segments in genome order,
    cuts on granule multiples, ranks balanced by TILES (20,000 scaffolds weigh their padded length)."""
    import random
    from cropsr_b200 import shard
    random.seed(1)
    for _ in range(3000):
        lengths = [random.choice([0, 1, 5, 127, 128, 129, 1000, 5000]) for _ in range(random.randint(0, 12))]
        world, granule = random.randint(1, 9), 128
        plans = shard.plan(lengths, world, granule)
        assert len(plans) == world
        pos = [0] * len(lengths)
        seen, last = set(), -1
        for k, a, b in (seg for segs in plans for seg in segs):
            assert k >= last and a == pos[k] and b <= lengths[k] and a % granule == 0
            last, pos[k] = k, b
            seen.add(k)
        assert pos == lengths and seen == set(range(len(lengths)))
        tiles = [sum(-(-(b - a) // granule) for _, a, b in p) for p in plans]
        assert max(tiles, default=0) <= -(-sum(-(-n // granule) for n in lengths) // world) + 2 + len(lengths)
    big = [80_000_000] * 6 + [random.randint(10_000, 500_000) for _ in range(4000)]
    tiles = [sum(-(-(b - a) // shard.TILE) for _, a, b in p) for p in shard.plan(big, 8)]
    assert max(tiles) - min(tiles) <= 2


def test_cli_help_and_missing_system(built_lib, capsys):
    """--help prints (argparse cannot wrap the auto-generated usage line with the reference's empty metavars);
    without --cas9 the CLI exits with the reference's message (CROPSR.py:335-336)."""
    from cropsr_b200 import cli
    with pytest.raises(SystemExit) as e:
        cli.main(["--help"])
    assert e.value.code == 0 and "--cas9" in capsys.readouterr().out
    with pytest.raises(SystemExit) as e:
        cli.main(["-f", "x.fa"])
    assert e.value.code == "Please select at least one CRISPR system: Cas9"
    assert cli.parse_devices("0-3") == [0, 1, 2, 3] and cli.parse_devices("0,2,5") == [0, 2, 5]


def test_candidate_scoring_source_on_the_cpu(tmp_path):
    """tests/host_emu.cu compiles the per-candidate functions of the scan kernel -- the SAME source,
    __host__ __device__ -- for the host and checks the table-driven Rule-Set-1 lanes against the dense
    replay of OpenBLAS' canonical lane order, the emit loop's score_hit against the generic
    extract_window + rs1_canonical pair, bit for bit (reference: CROPSR.py:285-313 through SURVEY 8c), and the
    PAM tests -- tile_hits, and the chunk counts k_pack writes into the tile headers -- against a byte-by-byte
    reading of random tokens with the bounds of CROPSR.py:415-430, for guides of 1 to 100,000 bases."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "host_emu")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-o", exe,
                    os.path.join(ROOT, "tests", "host_emu.cu")], check=True, capture_output=True, timeout=600)
    r = subprocess.run([exe, "300000", "300000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


def test_generated_rs1_tables_belong_to_the_weights_in_the_tree():
    """csrc/rs1_weights.inc is generated from cropsr_b200/rs1.py (minutes of hash search, so it is committed
    and not rebuilt on every make): it records the sha256 of the rs1.py it was made from."""
    import hashlib
    with open(os.path.join(ROOT, "cropsr_b200", "rs1.py"), "rb") as f:
        want = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(ROOT, "cropsr_b200", "csrc", "rs1_weights.inc")) as f:
        head = f.read(2000)
    m = re.search(r"rs1\.py sha256 ([0-9a-f]{64})", head)
    assert m and m.group(1) == want, "rs1.py changed: run `make -C cropsr_b200/csrc regen`"
