"""The CPU oracle against the reference's own known answers, the committed
golden vectors and (when oracle/_ref is present) the compiled reference itself.

These mirror reference tests/test_kmers.py:8-53 and pin oracle/kmers_oracle.c.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, REF_DATA
from oracle import oracle as orc_mod
from oracle.oracle import CompiledReference


# ---- reference tests/test_kmers.py:28-39 (known answers pin LSB-first encoding, Q2)
@pytest.mark.parametrize("kmer,value", [
    ("ATGCTAGCTAGAGAGAGAGGA", 696357446508),
    ("TTTTTTTTTTTTTTTTTTTTTTTTTTTT", 72057594037927935),
    ("TTTTTTTTTTTTTTTAGGCCCACTTTTT", 72006304612220927),
    ("AAAAAAAAAAAAAAAAAAAAAAAAAA", 0),
    ("GGGAGGGAGGGAGGGAGGGAGGGAGGG", 11868309606246954),
])
def test_kmer_to_int_reference_kat(oracle, kmer, value):
    assert oracle.kmer_to_int(kmer) == value


# ---- reference tests/test_kmers.py:42-53
@pytest.mark.parametrize("kmer,rc", [
    ("ATGCTAGCTAGAGAGAGAGGA", "TCCTCTCTCTCTAGCTAGCAT"),
    ("TTTTTTTTTTTTTTTTTTTTTTTTTTTT", "AAAAAAAAAAAAAAAAAAAAAAAAAAAA"),
    ("TTTTTTTTTTTTTTTAGGCCCACTTTTT", "AAAAAGTGGGCCTAAAAAAAAAAAAAAA"),
    ("AAAAAAAAAAAAAAAAAAAAAAAAAA", "TTTTTTTTTTTTTTTTTTTTTTTTTT"),
    ("GGGAGGGAGGGAGGGAGGGAGGGAGGG", "CCCTCCCTCCCTCCCTCCCTCCCTCCC"),
])
def test_reverse_complement_reference_kat(oracle, kmer, rc):
    assert oracle.reverse_complement(kmer) == rc


# ---- SURVEY 8a hash_function KATs (c/kmers.c:98-103)
@pytest.mark.parametrize("x,h", [
    (0, 0), (1, 1040391079), (696357446508, 1442241687), (2**42 - 1, 2460793505),
    (2**62 - 1, 2853814272), (2**64 - 1, 2825322496),
])
def test_hash_function_kat(oracle, x, h):
    assert oracle.hash_function(x) == h


def test_golden_kat(oracle, golden_kat):
    for s, v in golden_kat["kmer_to_int"]:
        assert oracle.kmer_to_int(s) == v, s
    for s, v in golden_kat["reverse_complement"]:
        assert oracle.reverse_complement(s) == v, s
    for x, v in golden_kat["hash_function"]:
        assert oracle.hash_function(x) == v, x


# ---- reference tests/test_kmers.py:8-11 and :14-25
def test_build_kmer_hash(oracle):
    s = oracle.set_from_file(os.path.join(REF_DATA, "hapA.txt"))
    assert s.contents.num_kmers == 4
    assert s.contents.k == 21
    assert s.contents.hash_size == 5


def test_count_kmers_in_read_fixture(oracle):
    a = oracle.set_from_file(os.path.join(REF_DATA, "hapA.txt"))
    b = oracle.set_from_file(os.path.join(REF_DATA, "hapB.txt"))
    read = "CTTATCATGTCTTTGTTTTCAAAGCTTCTTAGAGGTTTTTTTTTTTGGTGTTAATTGGCATAAATTATGGCT"
    assert oracle.count(read, a, b) == (2, 1)


def test_unfindable_kmer_q1(oracle, tmp_path):
    # a stored k-mer larger than its reverse complement is dead
    p = tmp_path / "a.txt"
    p.write_text("TTTTCAAAGCTTCTTAGAGGT\nACCTCTAAGAAGCTTTGAAAC\nACCTCTAAGAAGCTTTGAAAG\n")
    q = tmp_path / "b.txt"
    q.write_text("CCAGCTGTGTGATTGGTCCGC\nAAAAATAAGACAGCATGTGAC\nAACACCAAAAAAAAAAACCTC\n")
    a, b = oracle.set_from_file(str(p)), oracle.set_from_file(str(q))
    assert oracle.count("GGTTTTCAAAGCTTCTTAGAGGTCC", a, b) == (0, 0)
    assert oracle.count("GGACCTCTAAGAAGCTTTGAAAACC", a, b) == (0, 0)


def test_fixture_cli_counts(oracle, golden_cli):
    """reference tests/test_classify_by_kmers.py:19-57: the fixture run's TSV."""
    import gzip

    a = oracle.set_from_file(os.path.join(REF_DATA, "hapA.txt"))
    b = oracle.set_from_file(os.path.join(REF_DATA, "hapB.txt"))
    sa, sb = oracle.scaling_factors(a.contents.num_kmers, b.contents.num_kmers)
    assert (sa, sb) == (1.0, 4 / 3)
    lines = gzip.open(os.path.join(REF_DATA, "test.ccs.fastq.gz"), "rt").read().split("\n")
    out = []
    for i in range(0, len(lines) - 1, 4):
        name = lines[i][1:].split(" ")[0]
        ca, cb = oracle.count(lines[i + 1], a, b)
        x, y, bins = oracle.decide([ca], [cb], sa, sb)
        out.append("\t".join([name, chr(bins[0]), str(float(x[0])), str(float(y[0]))]))
    assert "\n".join(out) + "\n" == golden_cli["gz"]["stdout"]


def _load_case_sets(oracle, case):
    return (oracle.set_from_file(os.path.join(GOLDEN, case["set_a"])),
            oracle.set_from_file(os.path.join(GOLDEN, case["set_b"])))


def test_golden_counts(oracle, golden_counts):
    """Committed vectors from the compiled reference (unmodified, and zero-init
    scratch for reads with non-ACGT in their first k bases)."""
    for case in golden_counts:
        a, b = _load_case_sets(oracle, case)
        assert a.contents.num_kmers == case["num_kmers_a"]
        assert b.contents.num_kmers == case["num_kmers_b"]
        assert a.contents.k == case["k"]
        got = [list(oracle.count(r, a, b)) for r in case["reads"]]
        assert got == case["counts"], case["name"]
        oracle.set_free(a)
        oracle.set_free(b)


def test_streams_restatement_matches_oracle(oracle, golden_counts):
    """The flat F/E-stream + never-written-mask formulation used on the GPU."""
    for case in golden_counts:
        k = case["k"]
        keys = []
        for name in ("set_a", "set_b"):
            with open(os.path.join(GOLDEN, case[name])) as f:
                keys.append({oracle.kmer_to_int(line.rstrip("\n")) for line in f})
        for read, want in zip(case["reads"][:40], case["counts"][:40]):
            got = orc_mod.streams_count(read.encode(), keys[0], keys[1], k)
            assert list(got) == want, (case["name"], read[:60])


def test_batch_equals_single(oracle, golden_counts):
    case = golden_counts[0]
    a, b = _load_case_sets(oracle, case)
    reads = [r.encode() for r in case["reads"]]
    offs = np.zeros(len(reads) + 1, dtype=np.int64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8)
    ca, cb = oracle.count_batch(bases, offs, a, b)
    assert np.stack([ca, cb], 1).tolist() == case["counts"]


def test_degenerate_full_table_terminates(oracle, tmp_path):
    # Q6: num_kmers in {1,2} => hash_size == num_kmers => the reference spins forever on a miss
    p = tmp_path / "one.txt"
    p.write_text("ACGTACGTACGTACGTACGTA\n")
    s = oracle.set_from_file(str(p))
    assert s.contents.hash_size == 1
    assert oracle.count("ACGTACGTACGTACGTACGTAGGGG", s, s) == (1, 0)


def test_set_file_without_final_newline(oracle, tmp_path):
    p = tmp_path / "s.txt"
    p.write_text("ACCTCTAAGAAGCTTTGAAAA\nCTTATCATGTCTTTGTTTTCA\nCCAGCTGTGTGATTGGTCCGC")
    s = oracle.set_from_file(str(p))
    assert s.contents.num_kmers == 3 and s.contents.k == 21
    assert oracle.contains(s, oracle.kmer_to_int("CCAGCTGTGTGATTGGTCCGC"))


def test_decision_rule(oracle):
    sa, sb = oracle.scaling_factors(3, 7)
    assert sa == 7 / 3 and sb == 1.0
    a, b, bins = oracle.decide([3, 0, 0, 3, 7], [0, 2, 0, 7, 3], sa, sb)
    assert bytes(bins) == b"ABUUA"
    assert a[3] == 3 * (7 / 3) and b[3] == 7.0


# ---- live comparison with the compiled reference (build container; also on the GPU
# ---- box because oracle/_ref travels with the snapshot)
needs_ref = pytest.mark.skipif(not CompiledReference.available(), reason="oracle/_ref not built")


def _random_case(rng, k, n_kmers, n_reads, lo_invalid):
    g = "".join("ACGT"[i] for i in rng.integers(0, 4, 4000))
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}

    def pick():
        i = int(rng.integers(0, len(g) - k))
        s = g[i:i + k]
        if rng.random() < 0.5:
            s = "".join(comp[c] for c in reversed(s))
        return s
    A = [pick() for _ in range(n_kmers)]
    B = [pick() for _ in range(n_kmers)]
    reads = []
    for _ in range(n_reads):
        L = int(rng.integers(0, 700))
        i = int(rng.integers(0, len(g) - L))
        s = list(g[i:i + L])
        for pos in np.nonzero(rng.random(L) < 0.04)[0]:
            if pos >= lo_invalid:
                s[pos] = "NnacgtR"[int(rng.integers(0, 7))]
        reads.append("".join(s))
    return A, B, reads


@needs_ref
@pytest.mark.parametrize("k", [21, 31, 32, 13, 3])
def test_oracle_vs_compiled_reference_random(oracle, tmp_path, k):
    ref = CompiledReference()
    rng = np.random.default_rng(k)
    A, B, reads = _random_case(rng, k, 300, 150, lo_invalid=k)
    pa, pb = tmp_path / "a.txt", tmp_path / "b.txt"
    pa.write_text("".join(x + "\n" for x in A))
    pb.write_text("".join(x + "\n" for x in B))
    ra, rb = ref.set_from_file(str(pa)), ref.set_from_file(str(pb))
    oa, ob = oracle.set_from_file(str(pa)), oracle.set_from_file(str(pb))
    assert oa.contents.hash_size == ra.contents.hash_size
    for r in reads:
        assert oracle.count(r, oa, ob) == ref.count(r, ra, rb), r


@needs_ref
@pytest.mark.parametrize("k", [21, 32, 7])
def test_oracle_vs_zero_init_reference_ub_corner(oracle, tmp_path, k):
    ref = CompiledReference(zinit=True)
    rng = np.random.default_rng(100 + k)
    A, B, reads = _random_case(rng, k, 300, 150, lo_invalid=0)
    pa, pb = tmp_path / "a.txt", tmp_path / "b.txt"
    pa.write_text("".join(x + "\n" for x in A))
    pb.write_text("".join(x + "\n" for x in B))
    ra, rb = ref.set_from_file(str(pa)), ref.set_from_file(str(pb))
    oa, ob = oracle.set_from_file(str(pa)), oracle.set_from_file(str(pb))
    for r in reads:
        assert oracle.count(r, oa, ob) == ref.count(r, ra, rb), r


@needs_ref
def test_oracle_set_layout_equals_reference(oracle, golden_counts):
    """Same hash, same probe order => identical tables, slot for slot."""
    ref = CompiledReference()
    case = golden_counts[0]
    path = os.path.join(GOLDEN, case["set_a"])
    r, o = ref.set_from_file(path), oracle.set_from_file(path)
    n = r.contents.hash_size
    assert n == o.contents.hash_size
    rf = np.ctypeslib.as_array(r.contents.full, (n,))
    of = np.ctypeslib.as_array(o.contents.full, (n,))
    assert (rf == of).all()
    rk = np.ctypeslib.as_array(r.contents.kmers, (n,))[rf == 1]
    ok = np.ctypeslib.as_array(o.contents.kmers, (n,))[of == 1]
    assert (rk == ok).all()


@needs_ref
def test_ragged_set_file_q5(oracle, tmp_path):
    """Short / blank lines reuse stale bytes of the getline buffer (Q5)."""
    ref = CompiledReference()
    p = tmp_path / "ragged.txt"
    p.write_text("ACCTCTAAGAAGCTTTGAAAA\nCTTAT\n\nCCAGCTGTGTGATTGGTCCGC\nGG")
    r, o = ref.set_from_file(str(p)), oracle.set_from_file(str(p))
    n = r.contents.hash_size
    assert n == o.contents.hash_size and r.contents.num_kmers == o.contents.num_kmers == 5
    rf = np.ctypeslib.as_array(r.contents.full, (n,))
    rk = np.ctypeslib.as_array(r.contents.kmers, (n,))
    ok = np.ctypeslib.as_array(o.contents.kmers, (n,))
    assert (rk[rf == 1] == ok[rf == 1]).all()


# ---- the binary-search form of the oracle's set (full-size configs) equals the table form
@pytest.mark.parametrize("k", [9, 21, 25, 31])
def test_sorted_set_form_equals_table_form(oracle, k):
    from trio_binning_b200 import synth

    ka, kb = synth.sorted_disjoint_sets(3000, 2500, k, seed=k)
    assert np.all(np.diff(ka.astype(np.int64)) > 0) and np.all(np.diff(kb.astype(np.int64)) > 0)
    assert np.intersect1d(ka, kb).size == 0 and int(max(ka[-1], kb[-1])) < 4 ** k
    bases, offsets, _ = synth.random_reads(300, 400, k, ka, kb, seed=11, plant_every=60, invalid_rate=0.01)
    ta, tb = oracle.set_from_ints(ka, k), oracle.set_from_ints(kb, k)
    sa, sb = oracle.set_from_sorted(ka, k), oracle.set_from_sorted(kb, k)
    want = oracle.count_batch(bases, offsets, ta, tb)
    got = oracle.count_batch(bases, offsets, sa, sb)
    assert (want[0] == got[0]).all() and (want[1] == got[1]).all()
    assert want[0].sum() > 0 and want[1].sum() > 0
    thr = oracle.count_batch_threads(bases, offsets, sa, sb, threads=3)
    assert (want[0] == thr[0]).all() and (want[1] == thr[1]).all()
    for key in list(ka[:50]) + list(kb[:50]) + [0, 1, 4 ** k - 1]:
        assert oracle.contains(sa, int(key)) == oracle.contains(ta, int(key))
