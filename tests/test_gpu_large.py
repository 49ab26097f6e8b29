"""BASELINE.json configs[2..4] at their FULL set sizes, checked against the oracle.

    C3  150 M + 150 M 21-mers, long reads (mean 25 kb)   c/kmers.c:270-299 on configs[2]
    C4  400 M + 400 M 31-mers, long reads                configs[3]
    C5  50 M + 50 M 21-mers, 1 M reads of 150 bp         configs[4]
    plus k = 25 (a k nobody hand-tuned) and reads with 1 % non-ACGT bytes, which send about
    half of the 64-base spans through the kernels' general path (quirk Q3) at scale.

The checker is oracle/kmers_oracle.c itself -- the digit-space restatement of
count_kmers_in_read, sticky reverse-complement buffer included -- with its set in the
binary-search form (`orc_set_from_sorted`; pinned against the reference-shaped table form in
tests/test_oracle.py): membership is exact and independent of anything the GPU builds, and
needs no minutes-long table fill.  Sets are sorted by construction (synth.sorted_disjoint_sets).
Reads are a slice of each config's shape (sets are what make these configs large); every read
of the slice is compared for the long-read configs, every 20th (50 000 reads) for C5.
Bit-exact: counts and bins; scores are fp64 products and must be equal.
"""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = {
    # name: k, n_a, n_b, read sets [(n_reads, mean_len, fixed, invalid_rate, check_every)]
    "c3": (21, 150_000_000, 150_000_000, [(2000, 25_000, False, 0.0, 1), (1000, 25_000, False, 0.01, 1)]),
    "c4": (31, 400_000_000, 400_000_000, [(2000, 25_000, False, 0.0, 1), (500, 25_000, False, 0.01, 1)]),
    "c5": (21, 50_000_000, 50_000_000, [(1_000_000, 150, True, 0.0, 20), (200_000, 150, True, 0.01, 10)]),
    "k25": (25, 50_000_000, 50_000_000, [(1000, 20_000, False, 0.0, 1), (500, 20_000, False, 0.01, 1)]),
}


@pytest.fixture(scope="module")
def kmers():
    from trio_binning_b200 import kmers as mod

    assert mod.device_count() > 0, "GPU tests need a CUDA device"
    return mod


def _subsample(bases, offsets, every):
    if every == 1:
        return np.arange(len(offsets) - 1), bases, offsets
    idx = np.arange(0, len(offsets) - 1, every)
    lens = (offsets[idx + 1] - offsets[idx]).astype(np.int64)
    sub_off = np.zeros(len(idx) + 1, dtype=np.int64)
    np.cumsum(lens, out=sub_off[1:])
    if len(set(lens.tolist())) == 1:  # fixed-length reads: one fancy-index gather
        L = int(lens[0])
        sub = bases[(offsets[idx][:, None] + np.arange(L)[None, :]).ravel()]
    else:
        sub = np.concatenate([bases[offsets[i]:offsets[i + 1]] for i in idx])
    return idx, sub, sub_off


@pytest.mark.parametrize("name", list(CASES))
def test_full_size_sets_vs_oracle(kmers, oracle, name):
    from trio_binning_b200 import synth

    k, n_a, n_b, read_sets = CASES[name]
    t0 = time.time()
    ka, kb = synth.sorted_disjoint_sets(n_a, n_b, k, seed=1000 + k)
    assert len(ka) == n_a and len(kb) == n_b
    A = kmers.create_kmer_hash_set_from_ints(ka, k)
    B = kmers.create_kmer_hash_set_from_ints(kb, k)
    oa, ob = oracle.set_from_sorted(ka, k), oracle.set_from_sorted(kb, k)
    try:
        assert A.contents.n_distinct == n_a and B.contents.n_distinct == n_b
        assert kmers.get_number_kmers_in_set(A) == n_a
        fa, fb = oracle.scaling_factors(n_a, n_b)
        max_bases = max(int(n * (m if fx else m * 1.3)) + (1 << 20) for n, m, fx, _, _ in read_sets)
        max_reads = max(n for n, *_ in read_sets)
        clf = kmers.Classifier(A, B, max_bases, max_reads, 2)
        info = clf.filter_info()
        assert info["pair_buckets"] > 0 and info["minimizer"], "default kernel configuration expected"
        pb = kmers.PackedBatch(max_bases, max_reads)
        for si, (n_reads, mean_len, fixed, invalid, every) in enumerate(read_sets):
            bases, offsets, hap = synth.random_reads(n_reads, mean_len, k, ka, kb, seed=77 + si,
                                                     plant_every=100 if fixed else 1000,
                                                     fixed_len=fixed, invalid_rate=invalid)
            assert int(offsets[-1]) <= max_bases
            clf.submit(0, bases, offsets)
            ca, cb, sa, sb, bins = (x.copy() for x in clf.result(0))
            idx, sub, sub_off = _subsample(bases, offsets, every)
            wa, wb = oracle.count_batch_threads(sub, sub_off, oa, ob)
            bad = np.nonzero((ca[idx] != wa) | (cb[idx] != wb))[0]
            assert bad.size == 0, (name, si, "first mismatches", [(int(idx[i]), int(ca[idx[i]]), int(wa[i]),
                                                                   int(cb[idx[i]]), int(wb[i])) for i in bad[:5]])
            assert wa.sum() > 0 and wb.sum() > 0, "the sample must contain hits in both sets"
            # every read: score = count * scale in fp64 (equal, tolerance 0) and the A/B/U rule
            xa, xb, xbin = oracle.decide(ca, cb, fa, fb)
            assert (sa == xa).all() and (sb == xb).all() and (bins == xbin).all()
            # planted haplotype wins for nearly every read that got a plant it can find
            if invalid == 0.0 and not fixed:
                assert (bins[hap == 0] == ord("A")).mean() > 0.95 and (bins[hap == 1] == ord("B")).mean() > 0.95
            # host-packed route (2-bit batches over PCIe) gives the same answer
            pb.pack(bases, offsets, k)
            clf.submit_packed(1, pb)
            pa, pbb, *_ = clf.result(1)
            assert (pa == ca).all() and (pbb == cb).all()
        clf.close()
    finally:
        kmers.destroy_kmer_hash_set(A)
        kmers.destroy_kmer_hash_set(B)
        oracle.set_free(oa)
        oracle.set_free(ob)
    print(f"[{name}] {time.time() - t0:.1f} s", flush=True)
