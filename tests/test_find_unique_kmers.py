"""find-unique-kmers (reference src/trio_binning/find_unique_kmers.py).

CPU: the cut-off search against golden vectors generated from the REFERENCE's own
`analyze_histogram` (tests/golden/make_histogram_golden.py), the oracle's restatement of the
same, the command line, the key/text conversions.
GPU: counting, histogram, subtraction and dump against the numpy oracle
(oracle/find_unique_oracle.py, which restates KMC's documented behaviour -- KMC itself is not
available, so that part of the parity is unpinned and says so), the command-line tool end to end,
and the unique sets handed to the classifier with and without the text files in between."""
import contextlib
import gzip
import io
import json
import os
from unittest.mock import patch

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def fuk():
    from trio_binning_b200 import find_unique_kmers as mod

    return mod


@pytest.fixture(scope="module")
def golden_hist():
    with open(os.path.join(GOLDEN, "histogram_golden.json")) as f:
        return json.load(f)


def test_cutoff_search_equals_reference(fuk, golden_hist):
    """The reference's analyze_histogram (find_unique_kmers.py:106-168) on 79 histograms."""
    from oracle import find_unique_oracle as orc

    n_err = 0
    for case in golden_hist:
        lines = [tuple(x) for x in case["hist"]]
        with contextlib.redirect_stderr(io.StringIO()):
            try:
                got = list(fuk.analyze_histogram(lines, "h.txt"))
            except fuk.HistogramError as e:
                got = None
                assert "h.txt" in e.message
        assert got == case["result"], case["hist"][:8]
        try:
            want = list(orc.analyze_histogram(lines))
        except ValueError:
            want = None
        assert want == case["result"]
        n_err += got is None
    assert 0 < n_err < len(golden_hist)


def test_close_cutoffs_warn(fuk, capsys):
    lines = [(2, 90), (3, 10), (4, 20), (5, 30), (6, 5)]
    assert fuk.analyze_histogram(lines, "somewhere") == (3, 6)
    assert "not very far apart" in capsys.readouterr().err


def test_command_line_matches_the_reference(fuk):
    """reference find_unique_kmers.py:25-59: -k required, -p/-o/-s, --path-to-kmc, two positionals."""
    a = fuk.parse_args(["-k", "21", "a1.fq,a2.fq", "b.fq"])
    assert (a.kmer_size, a.threads, a.outpath, a.scratch_dir, a.path_to_kmc) == (21, 1, ".", ".", "kmc")
    assert a.read_files == ["a1.fq,a2.fq", "b.fq"]
    a = fuk.parse_args(["--kmer-size", "31", "-p", "8", "-o", "out", "-s", "tmp", "--path-to-kmc", "/x/kmc", "a", "b"])
    assert (a.kmer_size, a.threads, a.outpath, a.scratch_dir, a.path_to_kmc) == (31, 8, "out", "tmp", "/x/kmc")
    with pytest.raises(SystemExit):
        with contextlib.redirect_stderr(io.StringIO()):
            fuk.parse_args(["a", "b"])


def test_key_conversions(fuk, oracle):
    from oracle import find_unique_oracle as orc

    rng = np.random.default_rng(3)
    for k in (5, 21, 31, 32):
        seq = bytes(rng.choice(list(b"ACGT"), 300).astype(np.uint8))
        keys = np.unique(orc.canonical_kmers(seq, k))
        text = orc.keys_to_text(keys, k).decode().split()
        assert text == sorted(text)                                   # integer order = lexicographic order
        for key, s in zip(keys[:50].tolist(), text[:50]):
            rc = s[::-1].translate(str.maketrans("ACGT", "TGCA"))
            assert s <= rc and (s in seq.decode() or rc in seq.decode())
        lsb = fuk.keys_to_lsb_first(keys, k)
        assert [oracle.kmer_to_int(s) for s in text[:50]] == lsb[:50].tolist()   # what the classifier's loader computes


# ------------------------------------------------------------------------------------------ GPU
def _haplotypes(rng, length, snp_every):
    """Two haplotypes of one genome: identical except for a substitution about every snp_every bases."""
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    a = letters[rng.integers(0, 4, length)]
    b = a.copy()
    pos = np.sort(rng.choice(length, length // snp_every, replace=False))
    b[pos] = letters[(np.searchsorted(letters, a[pos]) + rng.integers(1, 4, len(pos))) % 4]
    return a, b


def _reads_from(rng, genome, coverage, read_len, err, n_rate=0.0, lower=False):
    n = int(len(genome) * coverage / read_len)
    comp = np.zeros(256, dtype=np.uint8)
    comp[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.frombuffer(b"TGCA", dtype=np.uint8)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    for _ in range(n):
        s = int(rng.integers(0, len(genome) - read_len))
        r = genome[s:s + read_len].copy()
        if rng.random() < 0.5:
            r = comp[r[::-1]]
        e = rng.random(read_len) < err
        r[e] = letters[rng.integers(0, 4, int(e.sum()))]
        if n_rate:
            r[rng.random(read_len) < n_rate] = ord("N")
        b = r.tobytes()
        if lower and rng.random() < 0.3:
            b = b.lower()
        out.append(b)
    return out


def _write_fastq(path, reads, gz=False):
    data = b"".join(b"@r%d\n" % i + r + b"\n+\n" + b"I" * len(r) + b"\n" for i, r in enumerate(reads))
    with (gzip.open(path, "wb", compresslevel=1) if gz else open(path, "wb")) as f:
        f.write(data)


@pytest.fixture(scope="module")
def kmers():
    from trio_binning_b200 import kmers as mod

    return mod


@pytest.mark.gpu
@pytest.mark.parametrize("k", [15, 21, 22, 23, 31, 32])   # 22 | 23: packed | split slots
def test_counting_histogram_and_subtraction_vs_oracle(kmers, k):
    from oracle import find_unique_oracle as orc

    rng = np.random.default_rng(100 + k if k != 22 else 222)   # a seed whose histogram has a usable valley
    ga, gb = _haplotypes(rng, 60_000, 300)
    reads_a = _reads_from(rng, ga, 25, 150, 0.004, n_rate=0.001, lower=True)
    reads_b = _reads_from(rng, gb, 18, 150, 0.004, n_rate=0.001)
    reads_a += [b"", b"ACGT", b"N" * 200, b"ACGTACGTAC" * 30]   # shorter than k, no valid base, low complexity
    ka, ca = orc.count_library(reads_a, k)
    kb, cb = orc.count_library(reads_b, k)
    with kmers.KmerCounter(k, 1 << 22) as A, kmers.KmerCounter(k, 1 << 22) as B:
        for i in range(0, len(reads_a), 1000):   # several batches
            A.add_reads([r.decode() for r in reads_a[i:i + 1000]])
        B.add_reads([r.decode() for r in reads_b])
        sa, sb = A.stats(), B.stats()
        assert sa["distinct"] == len(ka) and sa["windows"] == int(ca.sum())
        assert sb["distinct"] == len(kb) and sb["windows"] == int(cb.sum())
        ha = A.histogram(256)
        want = np.bincount(np.minimum(ca, 255), minlength=256)
        assert (ha == want).all()
        assert orc.histogram(ca) == [(c, int(ha[c])) for c in range(2, 256)]
        lo, hi = orc.analyze_histogram(orc.histogram(ca))
        lo_b, hi_b = orc.analyze_histogram(orc.histogram(cb))
        got_a = A.select_unique(B, lo, hi)
        got_b = B.select_unique(A, lo_b, hi_b)
        assert (got_a == orc.unique_kmers(ka, ca, kb, cb, lo, hi)).all() and len(got_a) > 100
        assert (got_b == orc.unique_kmers(kb, cb, ka, ca, lo_b, hi_b)).all() and len(got_b) > 100
        # every k-mer of the database, no subtraction: the whole dump
        everything = A.select_unique(None, 2, 255)
        assert (everything == ka[ca >= 2]).all()


@pytest.mark.gpu
def test_partitions_and_full_tables(kmers):
    from oracle import find_unique_oracle as orc

    k = 21
    rng = np.random.default_rng(5)
    ga, _ = _haplotypes(rng, 40_000, 500)
    reads = _reads_from(rng, ga, 10, 200, 0.002)
    keys, counts = orc.count_library(reads, k)
    texts = [r.decode() for r in reads]
    got = []
    for part in range(4):
        with kmers.KmerCounter(k, 1 << 17, part, 4) as c:
            c.add_reads(texts)
            got.append(c.select_unique(None, 1, 255))
            assert c.stats()["distinct"] < len(keys)
    # DB_MIN = 2 applies in select: singletons are not in a database
    assert (np.sort(np.concatenate(got)) == keys[counts >= 2]).all()
    with kmers.KmerCounter(k, 2048) as small:
        small.add_reads(texts)
        with pytest.raises(OverflowError):
            small.stats()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [21, 22, 25])
def test_counts_far_beyond_the_saturated_maximum(kmers, k):
    """A k-mer seen millions of times (homopolymers, satellites): up to k = 22 the count shares the
    slot's word with the key and stops near 2^19 -- it must never carry into nothing, corrupt the
    key, or leave the last histogram bin (readers clamp at 255, KMC's -cs)."""
    assert kmers.lib.tb_counter_slot_bytes(k) == (8 if k <= 22 else 12)
    with kmers.KmerCounter(k, 1 << 16) as c:
        for _ in range(3):
            c.add_reads(["A" * 400_000, "T" * 300_000, "ACGTTGCA" * 10])
        st = c.stats()
        assert st["windows"] == 3 * (400_000 - k + 1 + 300_000 - k + 1 + 80 - k + 1)
        hist = c.histogram(256)
        assert hist[255] == 1 and hist.sum() == st["distinct"]
        assert list(c.select_unique(None, 255, 255)) == [0]          # poly-A / poly-T: canonical key 0
        with kmers.KmerCounter(k, 1 << 16) as other:
            other.add_reads(["T" * 100] * 2)
            assert len(c.select_unique(other, 255, 255)) == 0          # and it is found by lookups


@pytest.mark.gpu
def test_find_unique_kmers_cli_and_classifier_handoff(kmers, fuk, oracle, tmp_path):
    """The command-line tool end to end (two files for haplotype A, one gzipped), its two output
    files against the oracle, then classify-by-kmers' sets built (a) from those text files and
    (b) straight from the keys: same sets, same classification."""
    from oracle import find_unique_oracle as orc
    from trio_binning_b200 import synth

    k = 21
    rng = np.random.default_rng(8)
    ga, gb = _haplotypes(rng, 80_000, 250)
    reads_a = _reads_from(rng, ga, 30, 150, 0.003)
    reads_b = _reads_from(rng, gb, 30, 150, 0.003)
    half = len(reads_a) // 2
    _write_fastq(tmp_path / "a1.fastq", reads_a[:half])
    _write_fastq(tmp_path / "a2.fastq.gz", reads_a[half:], gz=True)
    _write_fastq(tmp_path / "b.fastq", reads_b)
    out = tmp_path / "out"
    out.mkdir()
    argv = ["find-unique-kmers", "-k", str(k), "-o", str(out), "-s", str(out), "-p", "4",
            f"{tmp_path / 'a1.fastq'},{tmp_path / 'a2.fastq.gz'}", str(tmp_path / "b.fastq")]
    with patch("sys.argv", argv):
        fuk.main()
    ka, ca = orc.count_library(reads_a, k)
    kb, cb = orc.count_library(reads_b, k)
    ra = orc.analyze_histogram(orc.histogram(ca))
    rb = orc.analyze_histogram(orc.histogram(cb))
    want_a = orc.unique_kmers(ka, ca, kb, cb, *ra)
    want_b = orc.unique_kmers(kb, cb, ka, ca, *rb)
    assert (out / "hapA_only_kmers.txt").read_bytes() == orc.keys_to_text(want_a, k)
    assert (out / "hapB_only_kmers.txt").read_bytes() == orc.keys_to_text(want_b, k)
    assert len(want_a) > 200 and len(want_b) > 200
    assert [tuple(map(int, line.split())) for line in open(out / "haplotypeA.histogram")] == orc.histogram(ca)
    # hand-off: sets from the text files == sets from the keys
    A1 = kmers.create_kmer_hash_set(str(out / "hapA_only_kmers.txt"))
    B1 = kmers.create_kmer_hash_set(str(out / "hapB_only_kmers.txt"))
    keys_a, keys_b, got_ra, got_rb = fuk.find_unique_kmers([str(tmp_path / "a1.fastq"), str(tmp_path / "a2.fastq.gz")],
                                                           [str(tmp_path / "b.fastq")], k)
    assert (keys_a == want_a).all() and (keys_b == want_b).all() and got_ra == ra and got_rb == rb
    A2, B2 = fuk.classifier_sets(keys_a, keys_b, k)
    assert kmers.get_number_kmers_in_set(A1) == kmers.get_number_kmers_in_set(A2) == len(want_a)
    assert A1.contents.n_distinct == A2.contents.n_distinct
    test_reads = [ga[i:i + 5000].tobytes().decode() for i in range(0, 60_000, 7000)] + \
                 [gb[i:i + 5000].tobytes().decode() for i in range(0, 60_000, 7000)]
    r1 = kmers.classify_reads(test_reads, A1, B1)
    r2 = kmers.classify_reads(test_reads, A2, B2)
    assert r1 == r2
    n = len(test_reads) // 2
    assert all(r[4] == "A" for r in r1[:n]) and all(r[4] == "B" for r in r1[n:])
    # the oracle agrees with the classification (sets loaded from the same files)
    oa, ob = oracle.set_from_file(str(out / "hapA_only_kmers.txt")), oracle.set_from_file(str(out / "hapB_only_kmers.txt"))
    assert [(r[0], r[1]) for r in r1] == [oracle.count(t, oa, ob) for t in test_reads]
    # forced cut-offs and forced partitions give consistent answers
    k2a, k2b, _, _ = fuk.find_unique_kmers([str(tmp_path / "a1.fastq"), str(tmp_path / "a2.fastq.gz")],
                                           [str(tmp_path / "b.fastq")], k, cutoffs=[ra[0], ra[1], rb[0], rb[1]], n_parts=2)
    assert (k2a == want_a).all() and (k2b == want_b).all()
