"""Parity of the CUDA path (through the C ABI / the Python boundary) with the
oracle and the committed golden vectors.  Bit-exact: counts, bins; scores are
IEEE fp64 products and must be equal (tolerance stated where used: 1e-12 rel).

Mirrors reference tests/test_kmers.py and tests/test_classify_by_kmers.py, then
adds the cases the reference does not pin: non-ACGT handling,
unfindable k-mers, k != 21, ragged/duplicate set files, empty and tiny reads,
batch-shape independence, and full-size (BASELINE configs[1]) properties.
"""
import gzip
import hashlib
import os
from unittest.mock import patch

import numpy as np
import pytest

from conftest import GOLDEN, REF_DATA

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kmers():
    from trio_binning_b200 import kmers as mod

    assert mod.device_count() > 0, "GPU tests need a CUDA device"
    return mod


@pytest.fixture(scope="module")
def fixture_sets(kmers):
    a = kmers.create_kmer_hash_set(os.path.join(REF_DATA, "hapA.txt"))
    b = kmers.create_kmer_hash_set(os.path.join(REF_DATA, "hapB.txt"))
    return a, b


# ---------------------------------------------------------------- reference tests
def test_build_kmer_hash(kmers):  # reference tests/test_kmers.py:8-11
    s = kmers.create_kmer_hash_set(os.path.join(REF_DATA, "hapA.txt"))
    assert kmers.get_number_kmers_in_set(s) == 4
    assert s.contents.k == 21 and s.contents.hash_size == 5
    assert s.contents.n_distinct == 4


def test_count_kmers_in_read(kmers, fixture_sets):  # reference tests/test_kmers.py:14-25
    a, b = fixture_sets
    assert kmers.count_kmers_in_read(
        "CTTATCATGTCTTTGTTTTCAAAGCTTCTTAGAGGTTTTTTTTTTTGGTGTTAATTGGCATAAATTATGGCT", a, b) == (2, 1)


@pytest.mark.parametrize("kmer_str,kmer_int", [
    ("ATGCTAGCTAGAGAGAGAGGA", 696357446508),
    ("TTTTTTTTTTTTTTTTTTTTTTTTTTTT", 72057594037927935),
    ("TTTTTTTTTTTTTTTAGGCCCACTTTTT", 72006304612220927),
    ("AAAAAAAAAAAAAAAAAAAAAAAAAA", 0),
    ("GGGAGGGAGGGAGGGAGGGAGGGAGGG", 11868309606246954),
])
def test_kmer_to_int(kmers, kmer_str, kmer_int):  # reference tests/test_kmers.py:28-39
    assert kmers.kmer_to_int(kmer_str) == kmer_int


@pytest.mark.parametrize("kmer,revcomp_kmer", [
    ("ATGCTAGCTAGAGAGAGAGGA", "TCCTCTCTCTCTAGCTAGCAT"),
    ("TTTTTTTTTTTTTTTTTTTTTTTTTTTT", "AAAAAAAAAAAAAAAAAAAAAAAAAAAA"),
    ("TTTTTTTTTTTTTTTAGGCCCACTTTTT", "AAAAAGTGGGCCTAAAAAAAAAAAAAAA"),
    ("AAAAAAAAAAAAAAAAAAAAAAAAAA", "TTTTTTTTTTTTTTTTTTTTTTTTTT"),
    ("GGGAGGGAGGGAGGGAGGGAGGGAGGG", "CCCTCCCTCCCTCCCTCCCTCCCTCCC"),
])
def test_reverse_complement(kmers, kmer, revcomp_kmer):  # reference tests/test_kmers.py:42-53
    assert kmers.reverse_complement(kmer) == revcomp_kmer


def test_classify_by_kmers_help(capsys):  # reference tests/test_classify_by_kmers.py:10-16
    from trio_binning_b200.classify_by_kmers import main

    with patch("sys.argv", ["classify-by-kmers", "--help"]):
        with pytest.raises(SystemExit):
            main()
    out, _ = capsys.readouterr()
    assert "Classify reads into bins" in out


def _run_cli(tmpdir, reads, extra=()):
    from trio_binning_b200.classify_by_kmers import main

    argv = ["classify-by-kmers", reads, os.path.join(REF_DATA, "hapA.txt"), os.path.join(REF_DATA, "hapB.txt"),
            "--haplotype-a-out-prefix", os.path.join(tmpdir, "hapA"),
            "--haplotype-b-out-prefix", os.path.join(tmpdir, "hapB"),
            "--unclassified-out-prefix", os.path.join(tmpdir, "hapU")] + list(extra)
    with patch("sys.argv", argv):
        main()


def test_classify_by_kmers_cli_plain(capsys, tmpdir, golden_cli):
    """reference tests/test_classify_by_kmers.py:19-57, plus the TSV and the exact
    bytes of both files (the reference's overlap-bug artefact included)."""
    from trio_binning_b200.seq import readfq

    _run_cli(str(tmpdir), os.path.join(REF_DATA, "test.ccs.fastq.gz"), ["--no-gzip-output"])
    out, _ = capsys.readouterr()
    assert out == golden_cli["plain"]["stdout"]
    hap_a_out = readfq(open(os.path.join(tmpdir, "hapA.fastq"), "r"))
    correct_out = readfq(open(os.path.join(REF_DATA, "hapA.fastq"), "r"))
    num_reads = 0
    for seq1, seq2 in zip(hap_a_out, correct_out):
        num_reads += 1
        assert seq1.name == seq2.name
        assert seq1.seq == seq2.seq
        assert seq1.qual == seq2.qual
    assert num_reads == 1
    assert sorted(os.listdir(tmpdir)) == sorted(golden_cli["plain"]["files"])
    for name, meta in golden_cli["plain"]["files"].items():
        raw = open(os.path.join(tmpdir, name), "rb").read()
        assert hashlib.sha256(raw).hexdigest() == meta["sha256"], name
    assert open(os.path.join(tmpdir, "hapA.fastq"), "rb").read() == open(os.path.join(REF_DATA, "hapA.fastq"), "rb").read()


def test_classify_by_kmers_cli_gz(capsys, tmpdir, golden_cli):
    _run_cli(str(tmpdir), os.path.join(REF_DATA, "test.ccs.fastq.gz"))
    out, _ = capsys.readouterr()
    assert out == golden_cli["gz"]["stdout"]
    assert sorted(os.listdir(tmpdir)) == sorted(golden_cli["gz"]["files"])
    for name, meta in golden_cli["gz"]["files"].items():
        raw = gzip.decompress(open(os.path.join(tmpdir, name), "rb").read())
        assert hashlib.sha256(raw).hexdigest() == meta["sha256"], name


@pytest.mark.parametrize("name", ["test.fa", "test.fastq"])
def test_classify_by_kmers_cli_other_inputs(capsys, tmpdir, golden_cli, name):
    _run_cli(str(tmpdir), os.path.join(REF_DATA, name), ["--batch-bases", "50"])
    out, _ = capsys.readouterr()
    assert out == golden_cli[name]["stdout"]
    for fn, text in golden_cli[name]["files_text"].items():
        assert gzip.decompress(open(os.path.join(tmpdir, fn), "rb").read()).decode() == text


# ---------------------------------------------------------------- golden vectors
def test_golden_counts(kmers, golden_counts):
    """Vectors from the compiled reference: unmodified (non-ACGT only at positions
    >= k) and zero-initialised scratch (non-ACGT anywhere: the Q3-UB contract)."""
    for case in golden_counts:
        a = kmers.create_kmer_hash_set(os.path.join(GOLDEN, case["set_a"]))
        b = kmers.create_kmer_hash_set(os.path.join(GOLDEN, case["set_b"]))
        assert kmers.get_number_kmers_in_set(a) == case["num_kmers_a"]
        assert kmers.get_number_kmers_in_set(b) == case["num_kmers_b"]
        assert a.contents.k == case["k"]
        got = [list(c) for c in kmers.count_kmers_in_reads(case["reads"], a, b)]
        bad = [i for i, (g, w) in enumerate(zip(got, case["counts"])) if g != w]
        assert not bad, (case["name"], bad[:5], [got[i] for i in bad[:5]], [case["counts"][i] for i in bad[:5]])
        # the per-read legacy entry point gives the same numbers (batch of one)
        for i in range(0, len(case["reads"]), 7):
            assert list(kmers.count_kmers_in_read(case["reads"][i], a, b)) == case["counts"][i]
        kmers.destroy_kmer_hash_set(a)
        kmers.destroy_kmer_hash_set(b)


def test_golden_decisions_and_scores(kmers, oracle, golden_counts):
    case = golden_counts[0]
    a = kmers.create_kmer_hash_set(os.path.join(GOLDEN, case["set_a"]))
    b = kmers.create_kmer_hash_set(os.path.join(GOLDEN, case["set_b"]))
    res = kmers.classify_reads(case["reads"], a, b)
    sa, sb = oracle.scaling_factors(case["num_kmers_a"], case["num_kmers_b"])
    assert kmers.scaling_factors(a, b) == (sa, sb)  # fp64, bit-equal (<= 1e-12 rel required)
    counts = np.array(case["counts"], dtype=np.int32)
    wa, wb, wbin = oracle.decide(counts[:, 0], counts[:, 1], sa, sb)
    for i, (ca, cb, xa, xb, bn) in enumerate(res):
        assert [ca, cb] == case["counts"][i]
        assert xa == wa[i] and xb == wb[i]
        assert bn == chr(wbin[i])


# ---------------------------------------------------------------- randomised vs oracle
def _oracle_sets(oracle, keys_a, keys_b, k):
    return oracle.set_from_ints(keys_a, k), oracle.set_from_ints(keys_b, k)


@pytest.mark.parametrize("k,n_keys,mean_len,n_reads,fixed", [
    (21, 20000, 3000, 300, False),
    (31, 20000, 3000, 300, False),
    (32, 5000, 1000, 200, False),
    (21, 20000, 150, 4000, True),      # short-read shape (configs[4])
    (15, 3000, 500, 500, False),
    (1, 1, 200, 50, False),
    (2, 5, 200, 50, False),
])
def test_random_batches_vs_oracle(kmers, oracle, k, n_keys, mean_len, n_reads, fixed):
    from trio_binning_b200 import synth

    ka, kb = synth.random_kmer_sets(n_keys, n_keys // 2 + 1, k, seed=k * 7 + 1)
    bases, offsets, _ = synth.random_reads(n_reads, mean_len, k, ka, kb, seed=k, plant_every=200,
                                           fixed_len=fixed, invalid_rate=0.002)
    A = kmers.create_kmer_hash_set_from_ints(ka, k)
    B = kmers.create_kmer_hash_set_from_ints(kb, k)
    oa, ob = _oracle_sets(oracle, ka, kb, k)
    ca, cb, sa, sb, bins = kmers.classify_packed(bases, offsets, A, B)
    wa, wb = oracle.count_batch(bases, offsets, oa, ob)
    assert (ca == wa).all() and (cb == wb).all(), np.nonzero((ca != wa) | (cb != wb))[0][:10]
    if k > 2:
        assert wa.sum() > 0 and wb.sum() > 0  # the workload does exercise hits
    fa, fb = oracle.scaling_factors(len(ka), len(kb))
    xa, xb, xbin = oracle.decide(wa, wb, fa, fb)
    assert (sa == xa).all() and (sb == xb).all() and (bins == xbin).all()
    oracle.set_free(oa)
    oracle.set_free(ob)
    kmers.destroy_kmer_hash_set(A)
    kmers.destroy_kmer_hash_set(B)


def test_invalid_bases_everywhere_vs_oracle(kmers, oracle):
    """Non-ACGT bytes anywhere (incl. the first k bases -> fix-up kernel), long N runs,
    reads that are all N, lowercase, CR, and bytes >= 0x80."""
    from trio_binning_b200 import synth

    k = 21
    ka, kb = synth.random_kmer_sets(3000, 3000, k, seed=5)
    rng = np.random.default_rng(11)
    bases, offsets, _ = synth.random_reads(600, 400, k, ka, kb, seed=3, plant_every=60)
    bases = bases.copy()
    junk = np.frombuffer(b"NnacgtRY\r\x80\xff-", dtype=np.uint8)
    pos = rng.integers(0, len(bases), size=len(bases) // 25)
    bases[pos] = junk[rng.integers(0, len(junk), size=len(pos))]
    for r in range(0, 600, 9):  # N runs of assorted lengths at assorted places, some at read start
        s, e = offsets[r], offsets[r + 1]
        a = s if r % 2 else s + int(rng.integers(0, max(1, e - s)))
        bases[a:min(e, a + int(rng.integers(1, 120)))] = ord("N")
    bases[offsets[10]:offsets[11]] = ord("N")  # a read with no valid base at all
    A = kmers.create_kmer_hash_set_from_ints(ka, k)
    B = kmers.create_kmer_hash_set_from_ints(kb, k)
    oa, ob = _oracle_sets(oracle, ka, kb, k)
    ca, cb, *_ = kmers.classify_packed(bases, offsets, A, B)
    wa, wb = oracle.count_batch(bases, offsets, oa, ob)
    bad = np.nonzero((ca != wa) | (cb != wb))[0]
    assert len(bad) == 0, (bad[:10], ca[bad[:10]], wa[bad[:10]], cb[bad[:10]], wb[bad[:10]])
    assert wa.sum() > 100 and wb.sum() > 100


def test_edge_shapes(kmers, oracle, fixture_sets):
    a, b = fixture_sets
    oa = oracle.set_from_file(os.path.join(REF_DATA, "hapA.txt"))
    ob = oracle.set_from_file(os.path.join(REF_DATA, "hapB.txt"))
    hit = "CTTATCATGTCTTTGTTTTCAAAGCTTCTTAGAGGTTTTTTTTTTTGGTGTTAATTGGCATAAATTATGGCT"
    reads = ["", "A", "ACGT" * 5, hit[:21], hit[:20], hit, "", "", hit * 40, "N" * 100, hit[::-1], "",
             "n" + hit, hit + "N", "ACGTN" * 300]
    got = kmers.count_kmers_in_reads(reads, a, b)
    want = [oracle.count(r, oa, ob) for r in reads]
    assert got == want
    # empty batch, batch of only empty reads
    assert kmers.count_kmers_in_reads([], a, b) == []
    assert kmers.count_kmers_in_reads(["", ""], a, b) == [(0, 0), (0, 0)]
    # NUL byte cuts the read like strlen does (c/kmers.c:285); non-ASCII is UTF-8 bytes
    assert kmers.count_kmers_in_reads([hit + "\x00" + hit], a, b) == [oracle.count(hit, oa, ob)]
    assert kmers.count_kmers_in_reads(["é" + hit], a, b) == [oracle.count(("é" + hit).encode(), oa, ob)]


def test_batch_shape_independence(kmers):
    """Same reads, different batch compositions and orders -> same per-read result."""
    from trio_binning_b200 import synth

    k = 21
    ka, kb = synth.random_kmer_sets(5000, 5000, k, seed=9)
    bases, offsets, _ = synth.random_reads(257, 1500, k, ka, kb, seed=4, plant_every=100, invalid_rate=0.001)
    A = kmers.create_kmer_hash_set_from_ints(ka, k)
    B = kmers.create_kmer_hash_set_from_ints(kb, k)
    ca, cb, *_ = kmers.classify_packed(bases, offsets, A, B)
    reads = [bytes(bases[offsets[i]:offsets[i + 1]]).decode("latin-1") for i in range(257)]
    one_by_one = [kmers.count_kmers_in_read(r, A, B) for r in reads[:40]]
    assert one_by_one == list(zip(ca[:40].tolist(), cb[:40].tolist()))
    perm = np.random.default_rng(0).permutation(257)
    shuffled = kmers.count_kmers_in_reads([reads[i] for i in perm], A, B)
    assert shuffled == [(int(ca[i]), int(cb[i])) for i in perm]
    with kmers.Classifier(A, B, max_bases=int(offsets[-1]), max_reads=257, n_slots=2) as clf:
        half = 128
        clf.submit(0, bases[:offsets[half]], offsets[:half + 1])
        clf.submit(1, bases[offsets[half]:], offsets[half:] - offsets[half])
        r0, r1 = clf.result(0), clf.result(1)
        assert (np.concatenate([r0[0], r1[0]]) == ca).all()
        assert (np.concatenate([r0[1], r1[1]]) == cb).all()
        st = clf.stats(0)
        assert st["windows"] == synth.total_windows(offsets[:half + 1], k)
        assert st["probes"] == 2 * st["windows"] - int(ca[:half].sum())


# ---------------------------------------------------------------- set loader
def test_set_file_quirks(kmers, oracle, tmp_path):
    """Q1 (raw insert), Q5 (duplicates counted, ragged lines, missing final newline),
    k = 32 all-T (the table's empty sentinel), membership vs the oracle's table."""
    rng = np.random.default_rng(2)
    k = 32
    lines = ["".join("ACGT"[i] for i in rng.integers(0, 4, k)) for _ in range(500)]
    lines += lines[:20] + ["T" * 32, "A" * 32]
    p = tmp_path / "k32.txt"
    p.write_text("\n".join(lines))  # no final newline
    s = kmers.create_kmer_hash_set(str(p))
    o = oracle.set_from_file(str(p))
    assert kmers.get_number_kmers_in_set(s) == o.contents.num_kmers == 522
    assert s.contents.n_distinct + s.contents.has_all_ones == 502
    probe = np.array([oracle.kmer_to_int(x) for x in lines] + [int(x) for x in rng.integers(0, 2**63, 2000)],
                     dtype=np.uint64)
    want = np.array([oracle.contains(o, int(x)) for x in probe])
    assert (kmers.set_contains_ints(s, probe) == want).all()
    assert want[:522].all()

    ragged = tmp_path / "ragged.txt"
    ragged.write_text("ACCTCTAAGAAGCTTTGAAAA\nCTTAT\n\nCCAGCTGTGTGATTGGTCCGC\nGG")
    s2, o2 = kmers.create_kmer_hash_set(str(ragged)), oracle.set_from_file(str(ragged))
    assert kmers.get_number_kmers_in_set(s2) == o2.contents.num_kmers == 5
    n = o2.contents.hash_size
    stored = np.ctypeslib.as_array(o2.contents.kmers, (n,))[np.ctypeslib.as_array(o2.contents.full, (n,)) == 1]
    assert kmers.set_contains_ints(s2, stored).all()

    lower = tmp_path / "lower.txt"
    lower.write_text("acgtacgtacgtacgtacgta\nACGTNACGTACGTACGTACGT\n")
    s3, o3 = kmers.create_kmer_hash_set(str(lower)), oracle.set_from_file(str(lower))
    n = o3.contents.hash_size
    stored = np.ctypeslib.as_array(o3.contents.kmers, (n,))[np.ctypeslib.as_array(o3.contents.full, (n,)) == 1]
    assert kmers.set_contains_ints(s3, stored).all()

    for bad in ("", "\n", "A" * 40 + "\n"):
        q = tmp_path / "bad.txt"
        q.write_text(bad)
        with pytest.raises(RuntimeError):
            kmers.create_kmer_hash_set(str(q))


def test_degenerate_tiny_sets_terminate(kmers, tmp_path):
    # Q6: 1- and 2-line files make the reference spin forever on a miss; we terminate
    p = tmp_path / "one.txt"
    p.write_text("ACGTACGTACGTACGTACGTA\n")
    s = kmers.create_kmer_hash_set(str(p))
    assert kmers.count_kmers_in_read("ACGTACGTACGTACGTACGTAGGGGGGGGGGGGGGGGGGGGGGGGG", s, s) == (1, 0)


def test_mismatched_k_is_an_error(kmers, fixture_sets, tmp_path):
    p = tmp_path / "k5.txt"
    p.write_text("ACGTA\nCCCCC\nGGGGA\n")
    s5 = kmers.create_kmer_hash_set(str(p))
    with pytest.raises(RuntimeError, match="k differs"):
        kmers.count_kmers_in_reads(["ACGTACGTACGT"], fixture_sets[0], s5)


# ---------------------------------------------------------------- full size (configs[1])
def test_config2_full_size_properties(kmers, oracle, tmp_path_factory):
    """BASELINE configs[1]: 10k reads (mean 20 kb) vs 10M + 10M 21-mers, through the
    text loader.  Checked by (1) the oracle on a deterministic subsample of reads,
    (2) reverse-complement symmetry on all-ACGT reads (the multiset of canonical
    keys of a read and of its reverse complement are equal), (3) additivity:
    a read cut into pieces overlapping by k-1 bases has the counts of the whole."""
    from trio_binning_b200 import synth

    k, n = 21, 10_000_000
    tmp = str(tmp_path_factory.mktemp("c2"))
    ka, kb, pa, pb = synth.materialise_sets(tmp, n, n, k, seed=42)
    A = kmers.create_kmer_hash_set(pa)
    B = kmers.create_kmer_hash_set(pb)
    assert kmers.get_number_kmers_in_set(A) == n and A.contents.n_distinct == n
    bases, offsets, hap = synth.random_reads(10_000, 20_000, k, ka, kb, seed=43, plant_every=1000)
    ca, cb, sa, sb, bins = kmers.classify_packed(bases, offsets, A, B)
    # planted haplotype wins (findable plants only: about half of them, quirk Q1)
    assert (bins[hap == 0] == ord("A")).mean() > 0.97
    assert (bins[hap == 1] == ord("B")).mean() > 0.97
    # (1) oracle on every 50th read
    oa, ob = oracle.set_from_ints(ka, k), oracle.set_from_ints(kb, k)
    for r in range(0, 10_000, 50):
        assert oracle.count(bytes(bases[offsets[r]:offsets[r + 1]]), oa, ob) == (ca[r], cb[r]), r
    # (2) reverse complement of the whole batch, read by read
    comp = np.zeros(256, dtype=np.uint8)
    comp[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.frombuffer(b"TGCA", dtype=np.uint8)
    sub = slice(0, 2000)
    sub_off = offsets[: 2001]
    rc = np.empty(int(sub_off[-1]), dtype=np.uint8)
    for r in range(2000):
        rc[sub_off[r]:sub_off[r + 1]] = comp[bases[sub_off[r]:sub_off[r + 1]][::-1]]
    ra, rb, *_ = kmers.classify_packed(rc, sub_off, A, B)
    assert (ra == ca[sub]).all() and (rb == cb[sub]).all()
    # (3) additivity over chunks overlapping by k-1
    piece = 1000
    starts, lens, owner = [], [], []
    for r in range(500):
        s, e = int(offsets[r]), int(offsets[r + 1])
        for p in range(s, max(s + 1, e - (k - 1)), piece):
            starts.append(p)
            lens.append(min(e, p + piece + k - 1) - p)
            owner.append(r)
    poff = np.zeros(len(starts) + 1, dtype=np.int64)
    np.cumsum(lens, out=poff[1:])
    pbases = np.concatenate([bases[s:s + n_] for s, n_ in zip(starts, lens)])
    pa_, pb_, *_ = kmers.classify_packed(pbases, poff, A, B)
    owner = np.array(owner)
    assert (np.bincount(owner, weights=pa_, minlength=500) == ca[:500]).all()
    assert (np.bincount(owner, weights=pb_, minlength=500) == cb[:500]).all()


# ---------------------------------------------------------------- pre-filter on / off
@pytest.mark.parametrize("k", [21, 31, 32, 9])
def test_prefilter_does_not_change_results(kmers, oracle, k):
    """The L2-resident bitmap in front of the exact tables may only skip work: direct
    kernel (TB_FILTER_LOG2=0), a saturated 1-kbit bitmap (everything passes), a sparse one
    and the automatic size all give the oracle's counts."""
    from trio_binning_b200 import synth

    ka, kb = synth.random_kmer_sets(30000, 20000, k, seed=77 + k)
    bases, offsets, _ = synth.random_reads(400, 2500, k, ka, kb, seed=5, plant_every=150, invalid_rate=0.003)
    bases = bases.copy()
    bases[offsets[3]:offsets[3] + 5] = ord("N")          # exercise the fix-up path as well
    A = kmers.create_kmer_hash_set_from_ints(ka, k)
    B = kmers.create_kmer_hash_set_from_ints(kb, k)
    oa, ob = oracle.set_from_ints(ka, k), oracle.set_from_ints(kb, k)
    wa, wb = oracle.count_batch(bases, offsets, oa, ob)
    seen, pair_seen = set(), set()
    saved = {v: os.environ.get(v) for v in ("TB_FILTER_LOG2", "TB_PAIR", "TB_FILTER_BITS", "TB_MINIMIZER")}
    try:
        for pair, bits, mini in (("0", "1", "0"), ("1", "1", "0"), ("1", "2", "0"), ("1", "2", "1"), ("1", "1", "1")):
            os.environ["TB_PAIR"], os.environ["TB_FILTER_BITS"], os.environ["TB_MINIMIZER"] = pair, bits, mini
            for setting in ("0", "10", "20", "26", None):
                if setting is None:
                    os.environ.pop("TB_FILTER_LOG2", None)
                else:
                    os.environ["TB_FILTER_LOG2"] = setting
                with kmers.Classifier(A, B, int(offsets[-1]), len(offsets) - 1, 1) as clf:
                    info = clf.filter_info()
                    seen.add(info["log2_bits"])
                    pair_seen.add((info["pair_buckets"] > 0, info["bits_per_key"]))
                    clf.submit(0, bases, offsets)
                    ca, cb, *_ = clf.result(0)
                    assert (ca == wa).all() and (cb == wb).all(), (pair, bits, mini, setting)
                    if k <= 31 and pair == "1" and setting != "0":
                        assert info["minimizer"] == (mini == "1")
    finally:
        for v, val in saved.items():
            if val is None:
                os.environ.pop(v, None)
            else:
                os.environ[v] = val
    assert {0, 10, 20, 26} <= seen
    if k <= 31:
        assert {(False, 1), (True, 1), (True, 2)} <= pair_seen  # the pair table needs k <= 31
    else:
        assert pair_seen == {(False, 1)}


# ---------------------------------------------------------------- host-packed batches
@pytest.mark.parametrize("k", [21, 32, 6])
def test_host_packed_path_equals_ascii_path(kmers, oracle, k):
    """tb_pack_reads (host, AVX2) + tb_classifier_submit_packed vs the ASCII submit and the
    oracle, with non-ACGT bytes everywhere (dirty words, Q3-UB prefixes)."""
    from trio_binning_b200 import synth

    n_keys = 20000 if k > 8 else 900
    ka, kb = synth.random_kmer_sets(n_keys, n_keys, k, seed=300 + k)
    rng = np.random.default_rng(k)
    bases, offsets, _ = synth.random_reads(700, 900, k, ka, kb, seed=8, plant_every=80)
    bases = bases.copy()
    junk = np.frombuffer(b"NnacgtRY\r\x80-", dtype=np.uint8)
    pos = rng.integers(0, len(bases), size=len(bases) // 40)
    bases[pos] = junk[rng.integers(0, len(junk), size=len(pos))]
    for r in range(0, 700, 11):
        bases[offsets[r]:min(offsets[r + 1], offsets[r] + int(rng.integers(1, 80)))] = ord("N")
    A = kmers.create_kmer_hash_set_from_ints(ka, k)
    B = kmers.create_kmer_hash_set_from_ints(kb, k)
    oa, ob = oracle.set_from_ints(ka, k), oracle.set_from_ints(kb, k)
    wa, wb = oracle.count_batch(bases, offsets, oa, ob)
    total, n = int(offsets[-1]), len(offsets) - 1
    with kmers.Classifier(A, B, total, n, 2) as clf:
        pb = kmers.PackedBatch(total, n)
        pb.pack(bases, offsets, k)
        assert pb.c.n_ub > 0 and pb.c.n_dirty > 0
        clf.submit_packed(0, pb)
        clf.submit(1, bases, offsets)
        p = clf.result(0)
        a = clf.result(1)
        assert (p[0] == wa).all() and (p[1] == wb).all()
        assert (a[0] == wa).all() and (a[1] == wb).all()
        assert (p[2] == a[2]).all() and (p[3] == a[3]).all() and (p[4] == a[4]).all()
        # empty batch through the packed entry
        pb.pack(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64), k)
        clf.submit_packed(0, pb)
        assert len(clf.result(0)[0]) == 0


# ---------------------------------------------------------------- native pipeline vs Python I/O
def _write_fastq(path, bases, offsets, rng, width=0):
    import gzip as gz

    quals = bytes(range(33, 74))
    with (gz.open(path, "wb", compresslevel=1) if str(path).endswith(".gz") else open(path, "wb")) as f:
        for r in range(len(offsets) - 1):
            s = bytes(bases[offsets[r]:offsets[r + 1]])
            q = bytes(rng.choice(list(quals)) for _ in range(len(s)))
            if r % 7 == 3:  # some FASTA records in between
                f.write(b">read%d fasta\n%s\n" % (r, s))
            else:
                f.write(b"@read%d some comment\n%s\n+\n%s\n" % (r, s, q))


def test_native_pipeline_equals_python_io(kmers, capsys, tmp_path):
    """Same input through the native pipeline (default: 2-bit batches, one GPU, FASTQ-aware gzip
    encoder), the same with zlib only, on every visible GPU, on two, with three device workers, and
    through the Python reader/writer (--python-io) -- in gzip and in plain mode (the
    reference's --no-gzip-output with its two-handles artefact, seq.py:127-130): identical TSV and
    identical (decompressed) outputs, several batches."""
    import random

    from trio_binning_b200 import synth
    from trio_binning_b200.classify_by_kmers import main

    k = 21
    ka, kb = synth.random_kmer_sets(4000, 3000, k, seed=21)
    bases, offsets, _ = synth.random_reads(600, 1500, k, ka, kb, seed=22, plant_every=120, invalid_rate=0.002)
    pa, pb = tmp_path / "a.txt", tmp_path / "b.txt"
    synth.write_kmer_file(str(pa), ka, k)
    synth.write_kmer_file(str(pb), kb, k)
    reads = tmp_path / "reads.fastq.gz"
    _write_fastq(reads, bases, offsets, random.Random(3))
    for plain in (False, True):
        outs = {}
        for mode, extra in (("native", []), ("python", ["--python-io"]), ("native_zlib", ["--no-fastq-gzip"]),
                            ("native_all_gpus", ["--gpus", "0"]), ("native_2gpu", ["--gpus", "2"]),
                            ("native_3workers", ["--gpu-workers", "3"])):
            d = tmp_path / f"{mode}_{plain}"
            d.mkdir()
            argv = ["classify-by-kmers", str(reads), str(pa), str(pb), "--haplotype-a-out-prefix", str(d / "A"),
                    "--haplotype-b-out-prefix", str(d / "B"), "--unclassified-out-prefix", str(d / "U"),
                    "--batch-bases", "100000"] + extra + (["--no-gzip-output"] if plain else [])
            with patch("sys.argv", argv):
                main()
            tsv, _ = capsys.readouterr()
            files = {n: (open(d / n, "rb").read() if plain else gzip.decompress(open(d / n, "rb").read()))
                     for n in sorted(os.listdir(d))}
            outs[mode] = (tsv, files)
        assert sorted(outs["native"][1]) == (["A.fastq", "U.fastq"] if plain else ["A.fastq.gz", "B.fastq.gz", "U.fastq.gz"])
        for mode in outs:
            assert outs[mode][0] == outs["python"][0], (mode, plain)
            assert outs[mode][1] == outs["python"][1], (mode, plain)
        assert outs["native"][0].count("\n") == 600
        bins = [line.split("\t")[1] for line in outs["native"][0].splitlines()]
        assert {"A", "B", "U"} <= set(bins)


def test_set_clone_is_a_replica(kmers, oracle):
    """tb_set_clone (the per-GPU replicas of the one-process multi-GPU pipeline): same membership,
    same reference-visible fields, independent lifetime.  Cloned onto the last visible device (the
    same one on a single-GPU box)."""
    from trio_binning_b200 import synth

    k = 21
    ka, kb = synth.random_kmer_sets(50_000, 40_000, k, seed=31)
    A, B = kmers.create_kmer_hash_set_from_ints(ka, k), kmers.create_kmer_hash_set_from_ints(kb, k)
    dev = kmers.device_count() - 1
    A2, B2 = kmers.clone_kmer_hash_set(A, dev), kmers.clone_kmer_hash_set(B, dev)
    assert A2.contents.device == dev and A2.contents.num_kmers == len(ka) and A2.contents.k == k
    assert A2.contents.n_distinct == A.contents.n_distinct and A2.contents.d_table != A.contents.d_table
    bases, offsets, _ = synth.random_reads(300, 2000, k, ka, kb, seed=32, plant_every=150, invalid_rate=0.001)
    want = kmers.classify_packed(bases, offsets, A, B)
    kmers.destroy_kmer_hash_set(A)
    kmers.destroy_kmer_hash_set(B)
    got = kmers.classify_packed(bases, offsets, A2, B2)
    for w, g in zip(want, got):
        assert (w == g).all()
    oa, ob = _oracle_sets(oracle, ka, kb, k)
    wa, wb = oracle.count_batch(bases, offsets, oa, ob)
    assert (got[0] == wa).all() and (got[1] == wb).all()
    kmers.destroy_kmer_hash_set(A2)
    kmers.destroy_kmer_hash_set(B2)


# ---------------------------------------------------------------- the reference's own binding
def test_reference_style_ctypes_binding(kmers):
    """Bind the library exactly the way the reference's kmers.py does (5-field struct, the
    four symbols, c_char_p arguments; reference kmers.py:41-93) and replay its unit tests:
    this is the zero-edit drop-in of INTEGRATION.md section 1."""
    from ctypes import POINTER, Structure, byref, c_char_p, c_int, c_ubyte, c_uint64, cdll

    lib = cdll.LoadLibrary(kmers._lib_path)

    class RefHashSet(Structure):
        _fields_ = [("kmers", POINTER(c_uint64)), ("full", POINTER(c_ubyte)), ("hash_size", c_int),
                    ("k", c_ubyte), ("num_kmers", c_int)]

    lib.create_kmer_hash_set.argtypes = [c_char_p]
    lib.create_kmer_hash_set.restype = POINTER(RefHashSet)
    lib.count_kmers_in_read.argtypes = [c_char_p, POINTER(RefHashSet), POINTER(RefHashSet), POINTER(c_int), POINTER(c_int)]
    lib.kmer_to_int.argtypes = [c_char_p, c_ubyte]
    lib.kmer_to_int.restype = c_uint64
    lib.reverse_complement.argtypes = [c_char_p, c_char_p, c_ubyte]

    a = lib.create_kmer_hash_set(os.path.join(REF_DATA, "hapA.txt").encode("utf-8"))
    b = lib.create_kmer_hash_set(os.path.join(REF_DATA, "hapB.txt").encode("utf-8"))
    assert a.contents.num_kmers == 4 and b.contents.num_kmers == 3 and a.contents.k == 21
    ca, cb = c_int(), c_int()
    read = "CTTATCATGTCTTTGTTTTCAAAGCTTCTTAGAGGTTTTTTTTTTTGGTGTTAATTGGCATAAATTATGGCT"
    lib.count_kmers_in_read(read.encode("utf-8"), a, b, byref(ca), byref(cb))
    assert (ca.value, cb.value) == (2, 1)
    assert lib.kmer_to_int(b"ATGCTAGCTAGAGAGAGAGGA", c_ubyte(21)) == 696357446508
    out = bytes("x" * 21, "utf-8")
    lib.reverse_complement(b"ATGCTAGCTAGAGAGAGAGGA", out, 21)
    assert out.decode() == "TCCTCTCTCTCTAGCTAGCAT"
    # the reference driver's arithmetic on top of it (classify_by_kmers.py:72-76, :104-115)
    m = max(a.contents.num_kmers, b.contents.num_kmers)
    sa, sb = 1.0 * m / a.contents.num_kmers, 1.0 * m / b.contents.num_kmers
    assert (ca.value * sa, cb.value * sb) == (2.0, 4 / 3)


# ---------------------------------------------------------------- error behaviour
def test_api_misuse_fails_loudly(kmers, fixture_sets):
    a, b = fixture_sets
    bases = np.frombuffer(b"ACGT" * 100, dtype=np.uint8)
    offsets = np.array([0, 400], dtype=np.int64)
    with kmers.Classifier(a, b, max_bases=100, max_reads=4, n_slots=1) as clf:
        with pytest.raises(ValueError):
            clf.submit(0, bases, offsets)                      # larger than the slot
        rc = kmers.lib.tb_classifier_submit(clf._h, 3, None, None, 0, None, None, None, None, None)
        assert rc != 0 and b"slot" in kmers.lib.tb_last_error()
        rc = kmers.lib.tb_classifier_rerun_resident(clf._h, 0, 1, None, None)
        assert rc != 0 and b"no batch" in kmers.lib.tb_last_error()
    with pytest.raises(RuntimeError, match="non-decreasing|offsets"):
        kmers.classify_packed(bases, np.array([0, 300, 200, 400], dtype=np.int64), a, b)
    with pytest.raises(RuntimeError):
        kmers.create_kmer_hash_set_from_ints(np.zeros(0, dtype=np.uint64), 21)
    with pytest.raises(RuntimeError):
        kmers.create_kmer_hash_set_from_ints(np.arange(10, dtype=np.uint64), 40)
    with pytest.raises(RuntimeError):
        kmers.set_device(99)
    # a destroyed set is rejected instead of being dereferenced
    s = kmers.create_kmer_hash_set_from_ints(np.arange(10, dtype=np.uint64), 21)
    import ctypes as C

    shadow = C.cast(C.pointer(kmers._HashSet.from_buffer_copy(s.contents)), kmers.HashSet)
    kmers.destroy_kmer_hash_set(s)
    shadow.contents.magic = 0
    with pytest.raises(RuntimeError, match="invalid set"):
        kmers.classify_packed(bases, offsets, shadow, shadow)


# ---------------------------------------------------------------- short-read shape at volume
@pytest.mark.parametrize("k", [21, 31])
def test_short_read_volume_vs_oracle_subsample(kmers, oracle, k):
    """BASELINE configs[4] shape (150-bp reads, hundreds of thousands of them: every lane of the
    kernel straddles read boundaries) against 2 M + 2 M keys; the oracle checks every 100th read
    and the total of the counts must survive a shuffle of the batch."""
    from trio_binning_b200 import synth

    ka, kb = synth.random_kmer_sets(2_000_000, 2_000_000, k, seed=900 + k)
    n = 400_000
    bases, offsets, _ = synth.random_reads(n, 150, k, ka, kb, seed=k, plant_every=60, fixed_len=True,
                                           invalid_rate=0.0005)
    A = kmers.create_kmer_hash_set_from_ints(ka, k)
    B = kmers.create_kmer_hash_set_from_ints(kb, k)
    ca, cb, *_ = kmers.classify_packed(bases, offsets, A, B)
    oa, ob = oracle.set_from_ints(ka, k), oracle.set_from_ints(kb, k)
    for r in range(0, n, 100):
        assert oracle.count(bytes(bases[offsets[r]:offsets[r + 1]]), oa, ob) == (ca[r], cb[r]), r
    assert ca.sum() > 50_000 and cb.sum() > 50_000
    perm = np.random.default_rng(1).permutation(n)
    shuffled = bases.reshape(n, 150)[perm].reshape(-1)
    sa, sb, *_ = kmers.classify_packed(shuffled, offsets, A, B)
    assert (sa == ca[perm]).all() and (sb == cb[perm]).all()
    oracle.set_free(oa)
    oracle.set_free(ob)
    kmers.destroy_kmer_hash_set(A)
    kmers.destroy_kmer_hash_set(B)


# ---------------------------------------------------------------- round-2 additions
def test_kmer_in_hash_set_reference_signature(kmers, oracle, fixture_sets):
    """kmer_in_hash_set through the reference's own signature (c/kmers.c:245-268):
    char kmer_in_hash_set(char *kmer, char *empty_kmer_buffer, hash_set *set) -- canonical
    min(fwd, revcomp) lookup, the scratch buffer receives the reverse complement and keeps its
    old bytes where the k-mer holds a non-ACGT byte (quirk Q3)."""
    import ctypes as C

    a, b = fixture_sets
    lines_a = open(os.path.join(REF_DATA, "hapA.txt")).read().split()
    lines_b = open(os.path.join(REF_DATA, "hapB.txt")).read().split()
    oa = oracle.set_from_file(os.path.join(REF_DATA, "hapA.txt"))

    def call(kmer, s, scratch=b"x" * 21):
        buf = C.create_string_buffer(scratch, 22)
        got = kmers.lib.kmer_in_hash_set(kmer.encode(), buf, s)
        return got != b"\x00", buf.raw[:21].decode("latin-1")

    for line in lines_a:
        rc = kmers.reverse_complement(line)
        # every fixture k-mer is findable, in either orientation
        assert call(line, a) == (True, rc)
        assert call(rc, a) == (True, line)
        assert call(line, b)[0] is False
        key = min(oracle.kmer_to_int(line), oracle.kmer_to_int(rc))
        assert oracle.contains(oa, key)
    for line in lines_b:
        assert call(line, b)[0] is True and call(line, a)[0] is False
    # Q1: a stored k-mer larger than its reverse complement is unfindable in either orientation
    unfindable = "TTTTCAAAGCTTCTTAGAGGT"  # its reverse complement is hapA's line 1
    s = kmers.create_kmer_hash_set_from_ints(np.array([kmers.kmer_to_int(unfindable)], dtype=np.uint64), 21)
    assert call(unfindable, s)[0] is False and call(kmers.reverse_complement(unfindable), s)[0] is False
    kmers.destroy_kmer_hash_set(s)
    # Q3: a non-ACGT byte leaves the scratch slot as it was; forward code 0
    found, buf = call("ACGTNACGTACGTACGTACGT", a, scratch=b"." * 21)
    assert buf == kmers.reverse_complement("ACGTAACGTACGTACGTACGT")[:16] + "." + kmers.reverse_complement("ACGTAACGTACGTACGTACGT")[17:]
    assert found is False
    # agreement with the oracle on random k-mers, both strands, members and non-members
    from trio_binning_b200 import synth

    ka, _ = synth.random_kmer_sets(500, 1, 21, seed=3)
    S = kmers.create_kmer_hash_set_from_ints(ka, 21)
    oS = oracle.set_from_ints(ka, 21)
    probe = np.concatenate([ka[:200], synth.revcomp_ints(ka[200:400], 21), synth.random_kmer_sets(100, 1, 21, seed=4)[0]])
    n_found = 0
    for key in probe.tolist():
        text = synth.ints_to_ascii(np.array([key], dtype=np.uint64), 21).tobytes().decode()
        rc = oracle.reverse_complement(text)
        want = oracle.contains(oS, min(oracle.kmer_to_int(text), oracle.kmer_to_int(rc)))
        assert call(text, S) == (want, rc)
        n_found += want
    assert 50 < n_found < 400
    oracle.set_free(oS)
    kmers.destroy_kmer_hash_set(S)


def test_keys_outside_the_kmer_space_are_rejected(kmers):
    """a raw key with bits above 2k set would land in the pair table's tag bits."""
    good = np.array([1, 2, 3, (1 << 42) - 1], dtype=np.uint64)
    s = kmers.create_kmer_hash_set_from_ints(good, 21)
    assert s.contents.n_distinct == 4
    kmers.destroy_kmer_hash_set(s)
    for bad in (1 << 42, 1 << 62, (1 << 64) - 1):
        with pytest.raises(RuntimeError, match="4\\^k"):
            kmers.create_kmer_hash_set_from_ints(np.array([1, bad, 2], dtype=np.uint64), 21)
    s = kmers.create_kmer_hash_set_from_ints(np.array([(1 << 64) - 1, 5], dtype=np.uint64), 32)  # k=32: all 64 bits are digits
    assert s.contents.n_distinct == 1 and s.contents.has_all_ones == 1
    kmers.destroy_kmer_hash_set(s)


@pytest.mark.parametrize("k", list(range(11, 33)))
def test_every_k_vs_oracle(kmers, oracle, k):
    """Every k from 11 to 32: 13..31 run the unrolled phase 1 instantiated for that k, 11/12
    the runtime-k body, 32 the two-table kernel.  Clean and dirty (non-ACGT) spans, long and
    150-bp reads, text-file loader (so `k` really comes from the file's first line)."""
    import tempfile

    from trio_binning_b200 import synth

    ka, kb = synth.random_kmer_sets(30_000, 20_000, k, seed=100 + k)
    with tempfile.TemporaryDirectory() as tmp:
        pa, pb = os.path.join(tmp, "a.txt"), os.path.join(tmp, "b.txt")
        synth.write_kmer_file(pa, ka, k)
        synth.write_kmer_file(pb, kb, k)
        A, B = kmers.create_kmer_hash_set(pa), kmers.create_kmer_hash_set(pb)
    assert A.contents.k == k and kmers.get_number_kmers_in_set(A) == len(ka)
    oa, ob = _oracle_sets(oracle, ka, kb, k)
    for n_reads, mean_len, fixed, invalid in ((150, 4000, False, 0.0), (150, 3000, False, 0.01), (3000, 150, True, 0.001)):
        bases, offsets, _ = synth.random_reads(n_reads, mean_len, k, ka, kb, seed=k + n_reads, plant_every=100,
                                               fixed_len=fixed, invalid_rate=invalid)
        ca, cb, *_ = kmers.classify_packed(bases, offsets, A, B)
        wa, wb = oracle.count_batch(bases, offsets, oa, ob)
        assert (ca == wa).all() and (cb == wb).all(), (k, n_reads, np.nonzero((ca != wa) | (cb != wb))[0][:10])
        assert wa.sum() > 0 and wb.sum() > 0
    oracle.set_free(oa)
    oracle.set_free(ob)
    kmers.destroy_kmer_hash_set(A)
    kmers.destroy_kmer_hash_set(B)


def test_canonical_orientation_sets(kmers, oracle):
    """SURVEY 8(d) asks for both set generators: raw orientation (about half of the stored k-mers
    unfindable, Q1 -- every other synthetic test) and LSB-first-canonical orientation, where every
    stored k-mer is findable: twice the hits per planted k-mer, same kernels."""
    from trio_binning_b200 import synth

    k = 21
    ka, kb = synth.random_kmer_sets(200_000, 200_000, k, seed=8, canonical=True)
    assert (ka <= synth.revcomp_ints(ka, k)).all()
    bases, offsets, hap = synth.random_reads(2000, 5000, k, ka, kb, seed=9, plant_every=500)
    A, B = kmers.create_kmer_hash_set_from_ints(ka, k), kmers.create_kmer_hash_set_from_ints(kb, k)
    oa, ob = _oracle_sets(oracle, ka, kb, k)
    ca, cb, sa, sb, bins = kmers.classify_packed(bases, offsets, A, B)
    wa, wb = oracle.count_batch(bases, offsets, oa, ob)
    assert (ca == wa).all() and (cb == wb).all()
    # every planted k-mer is found (none overwritten by a later plant is rare): ~1 hit per 500 bases
    planted = int(((np.diff(offsets)[hap == 0] - k) // 500).clip(1).sum())
    assert ca[hap == 0].sum() > 0.9 * planted
    ra, rb = synth.random_kmer_sets(200_000, 200_000, k, seed=8, canonical=False)
    assert abs((ra <= synth.revcomp_ints(ra, k)).mean() - 0.5) < 0.01  # raw: half findable
    oracle.set_free(oa)
    oracle.set_free(ob)
    kmers.destroy_kmer_hash_set(A)
    kmers.destroy_kmer_hash_set(B)


def test_long_runs_of_non_acgt(kmers, oracle):
    """gap-filled / assembly-like input.  Runs of N from 1 kb to 400 kb -- inside reads,
    at their start (the Q3-UB fix-up path), at their end, spanning many tiles, whole reads of N --
    must stay exact on both routes (device pack and host pack), and must not take forever."""
    import time

    from trio_binning_b200 import synth

    k = 21
    ka, kb = synth.random_kmer_sets(50_000, 50_000, k, seed=41)
    ka = np.concatenate([ka, np.zeros(1, dtype=np.uint64)])  # key 0 (poly-A) is what an all-N window looks up
    rng = np.random.default_rng(42)
    bases, offsets, _ = synth.random_reads(40, 600_000, k, ka, kb, seed=43, plant_every=300)
    bases = bases.copy()
    n = len(offsets) - 1
    for r in range(n):
        s, e = int(offsets[r]), int(offsets[r + 1])
        length = int(rng.choice([1_000, 30_000, 200_000, 400_000]))
        length = min(length, e - s)
        where = r % 4
        a = s if where == 0 else (e - length if where == 1 else s + int(rng.integers(0, e - s - length + 1)))
        if where == 3:
            a, length = s, e - s  # the whole read
        bases[a:a + length] = ord("N")
        if r % 5 == 0:  # a second, short run and a lone lowercase base
            b2 = s + int(rng.integers(0, e - s - 50))
            bases[b2:b2 + 37] = ord("n")
    A, B = kmers.create_kmer_hash_set_from_ints(ka, k), kmers.create_kmer_hash_set_from_ints(kb, k)
    oa, ob = _oracle_sets(oracle, ka, kb, k)
    t0 = time.time()
    ca, cb, *_ = kmers.classify_packed(bases, offsets, A, B)
    t_gpu = time.time() - t0
    wa, wb = oracle.count_batch_threads(bases, offsets, oa, ob)
    bad = np.nonzero((ca != wa) | (cb != wb))[0]
    assert bad.size == 0, (bad[:5], ca[bad[:5]], wa[bad[:5]], cb[bad[:5]], wb[bad[:5]])
    assert wa.sum() > 100_000  # the all-N windows hit key 0 in set A by the thousand
    total = int(offsets[-1])
    with kmers.Classifier(A, B, total * 12, n, 1) as clf:
        # room for the Q3-UB prefixes the host packer ships for reads that start with a run (the
        # pipeline sends such a batch as ASCII when they do not fit: tb_pack_reads returns -3)
        pb = kmers.PackedBatch(total * 12, n)
        pb.pack(bases, offsets, k)
        clf.submit_packed(0, pb)
        pa, pbb, *_ = clf.result(0)
        assert (pa == wa).all() and (pbb == wb).all()
    assert t_gpu < 20, f"{t_gpu:.1f} s for {total} bases"
    oracle.set_free(oa)
    oracle.set_free(ob)
    kmers.destroy_kmer_hash_set(A)
    kmers.destroy_kmer_hash_set(B)


@pytest.mark.parametrize("k", [20, 21, 22, 23, 24])
def test_two_level_bitmap_addressing(kmers, oracle, k):
    """The two-level addressing experiment (TB_HIER=1): bitmaps of >= 2^31 bits at
    k = 20..24 address their 128-byte lines by a second, coarser minimizer.  Forced here on small sets
    with TB_FILTER_LOG2=31.  Clean reads take the unrolled path, reads with non-ACGT bytes the general
    one; both must agree with the oracle, and with the one-level result."""
    from trio_binning_b200 import synth

    ka, kb = synth.random_kmer_sets(300_000, 200_000, k, seed=300 + k)
    A, B = kmers.create_kmer_hash_set_from_ints(ka, k), kmers.create_kmer_hash_set_from_ints(kb, k)
    oa, ob = _oracle_sets(oracle, ka, kb, k)
    batches = [synth.random_reads(300, 6000, k, ka, kb, seed=k, plant_every=150)[:2],
               synth.random_reads(300, 5000, k, ka, kb, seed=k + 1, plant_every=150, invalid_rate=0.01)[:2],
               synth.random_reads(20_000, 150, k, ka, kb, seed=k + 2, plant_every=80, fixed_len=True)[:2]]
    results = {}
    saved = {name: os.environ.get(name) for name in ("TB_FILTER_LOG2", "TB_HIER")}
    try:
        for hier in ("1", "0"):
            os.environ["TB_FILTER_LOG2"] = "31"
            os.environ["TB_HIER"] = hier
            total = max(int(o[-1]) for _, o in batches)
            with kmers.Classifier(A, B, total, 20_000, 1) as clf:
                info = clf.filter_info()
                assert info["log2_bits"] == 31 and info["two_level"] == (hier == "1")
                for i, (bases, offsets) in enumerate(batches):
                    clf.submit(0, bases, offsets)
                    ca, cb, *_ = clf.result(0)
                    results[(hier, i)] = (ca.copy(), cb.copy())
                    assert clf.sector_runs(0) > 0
    finally:
        for name, value in saved.items():
            if value is None:
                os.environ.pop(name, None)
            else:
                os.environ[name] = value
    for i, (bases, offsets) in enumerate(batches):
        wa, wb = oracle.count_batch(bases, offsets, oa, ob)
        for hier in ("1", "0"):
            ca, cb = results[(hier, i)]
            assert (ca == wa).all() and (cb == wb).all(), (k, hier, i)
        assert wa.sum() > 0 and wb.sum() > 0
    oracle.set_free(oa)
    oracle.set_free(ob)
    kmers.destroy_kmer_hash_set(A)
    kmers.destroy_kmer_hash_set(B)


@pytest.mark.gpu
def test_measurement_aids_run(kmers):
    """bench.py's denominators come from the library itself: the random-sector gather and the plain
    host-to-device copy.  Smoke only: they run and report a plausible rate."""
    rate, ms = kmers.gather_bench(64 << 20, 8, 256, 8, 8, 0)
    assert rate > 1e9 and ms > 0
    assert kmers.h2d_bench(16 << 20, 8) > 1.0   # GB/s
